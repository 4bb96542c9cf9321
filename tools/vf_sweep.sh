#!/bin/bash
# A/B of the very-far Jacobian kernel's geometry on one configs[4] path: rebuilds lbl_jac.o on the GPU box per variant
cd arts_b200/csrc
for cfg in "4 5 1 2" "4 6 1 2" "4 6 2 2" "4 5 1 4" "5 5 1 4" "5 5 2 4" "4 5 2 2"; do
  set -- $cfg
  rm -f lbl_jac.o
  make EXTRA="-DVF_MB4=$1 -DVF_MB2=$2 -DVF_UNROLL=$3" > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  grep -A2 "vfar_kernelILi2ELi$4" lbl_jac.ptxas.log | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../..; echo "cfg MB4=$1 MB2=$2 UNROLL=$3 R=$4: $(AB200_JAC_VFAR_R=$4 python tools/c5_jac_probe.py --reps 5 | python -c 'import json,sys; d=json.load(sys.stdin); print(d["T+VMR"]["propmat_ms"], d["T+3VMR"]["propmat_ms"])')")
done
rm -f lbl_jac.o; make > /dev/null 2>&1
