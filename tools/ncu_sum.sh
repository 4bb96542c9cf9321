#!/bin/bash
# full ncu capture of the line-sum kernel on the reduced-level bench shape (1 GPU)
set -x
CMD="python bench.py --steps 2 --warmup 3 --levels 10 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lbl_sum_real -s 3 -c 1 -o gpurun_out/prof_sum -f $CMD > gpurun_out/ncu_sum.log 2>&1
tail -2 gpurun_out/ncu_sum.log | cut -c1-300
