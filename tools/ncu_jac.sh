#!/bin/bash
# ncu capture of the Jacobian line-sum kernel on one config-5 path (run only after the plain command exited 0)
set -e
mkdir -p gpurun_out
python tools/c5_jac_probe.py --reps 3 > gpurun_out/c5_jac_probe.json
cat gpurun_out/c5_jac_probe.json
ncu --set full --clock-control none --import-source on -k regex:lbl_sum_jac -c 2 -o gpurun_out/jac python tools/c5_jac_probe.py --reps 1 > gpurun_out/ncu_jac.log 2>&1 || tail -5 gpurun_out/ncu_jac.log
ncu -i gpurun_out/jac.ncu-rep --page details --csv > gpurun_out/jac_details.csv 2>/dev/null || true
ls -la gpurun_out/jac.ncu-rep
