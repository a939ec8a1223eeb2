#!/bin/bash
# DRAM traffic (dram__bytes_read/write.sum) of every launch of the dominant kernel in ONE step -> gpurun_out/<tag>_traffic.csv
TAG=${1:-traffic}
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'tc_gemm_ts_kernel' -s 177 -c 59 --csv --log-file gpurun_out/${TAG}_traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_traffic.log 2>&1
tail -2 gpurun_out/${TAG}_traffic.log
wc -l gpurun_out/${TAG}_traffic.csv
