#!/bin/bash
# Multi-GPU bench on N GPUs of one box.  usage: bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-mg}
N=${2:-2}
mkdir -p gpurun_out
export REFIL_BENCH_DEBUG=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
grep -E "^rank|Error|error" gpurun_out/${TAG}_bench_n${N}.err | tail -20
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, d["e2e"]["value"], d["clocks"])
print("env", d["env"]["value"])
PY
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
PY
