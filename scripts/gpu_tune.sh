#!/bin/bash
# A/B of the submission knobs on one GPU: grouped hypernetworks on/off, m-tiles per CTA, GRU scan variant.
# usage: bash scripts/gpu_tune.sh <tag> "<batches>" "<labels>"
TAG=${1:-tune}
BATCHES=${2:-"16 64 128"}
LABELS=${3:-"group nogroup gru_mma"}
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
for B in $BATCHES; do
  for L in $LABELS; do
    case $L in
      group) EXTRA="" run group $B A=1;;
      nogroup) EXTRA="--no-group" run nogroup $B A=1;;
      gru_mma) EXTRA="" run gru_mma $B REFIL_GRU_MODE=mma;;
      crit*) EXTRA="--crit-tiles ${L#crit}" run $L $B A=1;;
      noprio*) EXTRA="--no-prio --crit-tiles ${L#noprio}" run $L $B A=1;;
      attn_generic) EXTRA="" run attn_generic $B REFIL_ATTN=generic;;
      y_ldg) EXTRA="" run y_ldg $B REFIL_TCW_Y=ldg;;
      pdl0) EXTRA="" run pdl0 $B REFIL_PDL=0;;
      pdl0_nogroup) EXTRA="--no-group" run pdl0_nogroup $B REFIL_PDL=0;;
      tiles*) EXTRA="" run $L $B REFIL_TC_MIN_TILES=${L#tiles};;
      chunks*) EXTRA="" run $L $B REFIL_TC_MIN_CHUNKS=${L#chunks};;
    esac
  done
done
