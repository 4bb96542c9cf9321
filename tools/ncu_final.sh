#!/bin/bash
# ncu evidence of the current kernels (1 GPU): launch list of the bench shape + full captures
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lbl_sum_real -s 3 -c 1 -o gpurun_out/prof_sum -f $CMD > gpurun_out/ncu_sum.log 2>&1
$CMD > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stokes_chain -s 3 -c 1 -o gpurun_out/prof_stokes -f $CMD > gpurun_out/ncu_stokes.log 2>&1
tail -c 300 gpurun_out/ncu_plain.log; ls -la gpurun_out | tail -8
