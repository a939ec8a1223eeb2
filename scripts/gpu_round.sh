#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one `--set full` capture of the top kernels.
# usage (on the GPU box, from the repo root): bash scripts/gpu_round.sh <tag> [tests|notests] [full|nofull]
TAG=${1:-rXX}
TESTS=${2:-tests}
FULL=${3:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv 2>/dev/null &
SMI=$!
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<EOF
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"])
print(d["kernels_ms_per_step"])
for k in ("roofline", "roofline_attention", "roofline_wgrad", "roofline_attention_bwd"):
    if d.get(k): print(k, round(d[k]["frac"], 3), round(d[k]["ms_per_step"], 3))
print("env", d["env"]["value"])
EOF
kill $SMI 2>/dev/null
# launch list of one short run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_ncu_launches.log 2>&1
if [ "$FULL" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'tc_gemm_t._kernel|tc_wgrad_ts_kernel|tc_gemm_wgrad_kernel|attn_fwd_kernel|attn_bwd_kernel|gru_scan' -s 60 -c 16 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/
fi
