#!/bin/bash
# ncu --set full of the dense kernels on the in_trans shape.  usage: bash scripts/gpu_prof.sh <tag> [kernel regex] [script]
TAG=${1:-prof}
REGEX=${2:-tc_gemm}
SCRIPT=${3:-scripts/prof_dense.py}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 3 -c 3 -o gpurun_out/${TAG}_prof -f \
  python $SCRIPT > gpurun_out/${TAG}_ncu.log 2>&1
tail -5 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
