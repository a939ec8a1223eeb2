#!/bin/bash
# Per-rank view of strong scaling on ONE GPU: the step a rank runs when the named global batch is sharded over 8 / 4 / 2 GPUs
# (B = 16 / 32 / 64 episodes), without the 1.65 MB all-reduce.  usage: bash scripts/gpu_strong.sh <tag>
TAG=${1:-strong}
mkdir -p gpurun_out
for W in ns-strong cfg5; do
  for B in 16 32 64; do
    timeout 300 python bench.py --workload $W --batch $B --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_${W}_b${B}.json 2> gpurun_out/${TAG}_${W}_b${B}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_${W}_b${B}.json"))
print("$W B=$B", round(d["value"]), "tr/s", round(d["ms_per_step"], 3), "ms/step", d["launches_per_step"], "launches; serialised kernels:",
      round(sum(d["kernels_ms_per_step"].values()), 3), "ms")
print("   ", d["kernels_ms_per_step"])
PY
  done
done
