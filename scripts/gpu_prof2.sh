#!/bin/bash
# usage: bash scripts/gpu_prof2.sh <tag> <kernel regex> <skip> <count> <script>
TAG=$1; REGEX=$2; SKIP=$3; COUNT=$4; SCRIPT=$5
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -o gpurun_out/${TAG}_prof -f \
  python $SCRIPT > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
