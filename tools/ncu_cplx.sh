#!/bin/bash
# ncu capture of the complex line-sum kernel on config 3 (run only after the plain command exited 0)
set -e
mkdir -p gpurun_out
python tools/c3_probe.py --reps 3 > gpurun_out/c3_probe.json
cat gpurun_out/c3_probe.json
ncu --set full --clock-control none --import-source on -k regex:lbl_sum_cplx -c 1 -o gpurun_out/cplx python tools/c3_probe.py --reps 1 > gpurun_out/ncu_cplx.log 2>&1 || tail -5 gpurun_out/ncu_cplx.log
ls -la gpurun_out/cplx.ncu-rep
