#!/bin/bash
# ncu --set full with source correlation of the attention backward (and forward) kernels of one eager step.
# usage: bash scripts/gpu_attn_prof.sh <tag> [kernel regex] [launch count]
TAG=${1:-attnprof}
KRE=${2:-attn_bwd_h4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -c ${3:-6} \
  -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source_cuda.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
if [ $(stat -c %s gpurun_out/${TAG}_prof.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/${TAG}_prof.ncu-rep; fi
