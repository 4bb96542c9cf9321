#!/bin/bash
# first GPU contact: parity tests, then a tiny timing probe
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
