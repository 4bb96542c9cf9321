#!/bin/bash
# full ncu capture of the far-field line-sum kernels (prepare, moments, scan, far, near) on one level batch of the bench shape
CMD="python bench.py --steps 1 --warmup 2 --levels 33 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lbl_fmm|lbl_prepare" -s 12 -c 6 -o gpurun_out/${1:-r2t}_fmm -f $CMD > gpurun_out/ncu_fmm.log 2>&1
tail -2 gpurun_out/ncu_fmm.log | cut -c1-300
