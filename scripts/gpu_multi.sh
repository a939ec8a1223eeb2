#!/bin/bash
# Multi-GPU run on N GPUs of one box: the NCCL parity tests (2 ranks) and the bench line at N (weak headline + strong record).
# usage: bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-mg}
N=${2:-2}
mkdir -p gpurun_out
export REFIL_BENCH_DEBUG=1
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -x -q > gpurun_out/${TAG}_pytest_n${N}.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_n${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
grep -E "Error|error" gpurun_out/${TAG}_bench_n${N}.err | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "launches_per_step")}, d["e2e"]["value"], d["clocks"])
print("strong", {k: (v.get("value"), v.get("ms_per_step"), v.get("episodes_per_gpu")) for k, v in d["strong"].items()})
print("env", d["env"]["value"], d["env"].get("rollout_cfg4"))
PY
