#!/bin/bash
# A/B of the submission knobs on one GPU: grouped hypernetworks on/off, m-tiles per CTA, chunks per weight-gradient split.
# usage: bash scripts/gpu_tune.sh <tag>
TAG=${1:-tune}
run() {  # label, B, env..., extra args
  local label=$1 B=$2; shift 2
  local out=gpurun_out/${TAG}_${label}_b${B}.json
  env "$@" timeout 300 python bench.py --workload ns-strong --batch $B --steps 20 --warmup 5 --no-cpu --no-extra $EXTRA > $out 2>/dev/null
  python - <<PY
import json
try:
    d = json.load(open("$out"))
    print("%-28s B=%-4d %8.0f tr/s %7.3f ms/step %5.1f launches  serialised %.3f ms" % ("$label", $B, d["value"], d["ms_per_step"], d["launches_per_step"], sum(d["kernels_ms_per_step"].values())))
except Exception as e:
    print("$label B=$B failed", e)
PY
}
for B in 16 128; do
  EXTRA="" run group $B A=1
  EXTRA="--no-group" run nogroup $B A=1
  EXTRA="" run tiles2 $B REFIL_TC_MIN_TILES=2
  EXTRA="" run tiles3 $B REFIL_TC_MIN_TILES=3
  EXTRA="" run tiles12 $B REFIL_TC_MIN_TILES=12
  EXTRA="" run chunks8 $B REFIL_TC_MIN_CHUNKS=8
  EXTRA="" run chunks64 $B REFIL_TC_MIN_CHUNKS=64
  EXTRA="" run gru_mma $B REFIL_GRU_MODE=mma
done
