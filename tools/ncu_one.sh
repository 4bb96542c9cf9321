#!/bin/bash
# One ncu pass per call over the bench shape (1 GPU), each only after the same command exited 0 without ncu.
#   tools/ncu_one.sh launches | sum | stokes
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_$1.log 2>&1
case "$1" in
  launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1 ;;
  sum)      ncu --set full --clock-control none --import-source on -k regex:lbl_sum_real -s 3 -c 1 -o gpurun_out/prof_sum -f $CMD > gpurun_out/ncu_sum.log 2>&1 ;;
  stokes)   ncu --set full --clock-control none --import-source on -k regex:stokes_chain -s 3 -c 1 -o gpurun_out/prof_stokes -f $CMD > gpurun_out/ncu_stokes.log 2>&1 ;;
esac
tail -c 200 gpurun_out/ncu_plain_$1.log; ls -la gpurun_out | tail -5
