#!/bin/bash
# Development iteration on the GPU box: dense-kernel tests first (fast fail), then the whole GPU suite, micro-benchmarks
# and the bench line.  usage: bash scripts/gpu_iter.sh <tag> [extra pytest -k filter for the first stage]
TAG=${1:-it}
FILTER=${2:-tc_}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$FILTER" > gpurun_out/${TAG}_pytest_k.log 2>&1
RC=$?
tail -15 gpurun_out/${TAG}_pytest_k.log
if [ $RC -ne 0 ]; then
  echo "kernel tests failed (rc=$RC): compute-sanitizer on the first failing stage"
  timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$FILTER" > gpurun_out/${TAG}_sanitizer.log 2>&1
  grep -v "^$" gpurun_out/${TAG}_sanitizer.log | head -60
  exit 1
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/${TAG}_bench_gemm.log 2>&1
cat gpurun_out/${TAG}_bench_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 800 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["clocks"])
print(d["kernels_ms_per_step"])
print(d.get("dense_shapes_us"))
for k in ("roofline", "roofline_attention", "roofline_wgrad", "roofline_attention_bwd"):
    if d.get(k): print(k, round(d[k]["frac"], 3), round(d[k]["ms_per_step"], 3))
print("env", d["env"]["value"])
PY
