#!/bin/bash
# compute-sanitizer memcheck over the kernel-level GPU tests.  usage: bash scripts/gpu_sanitize.sh <tag> [pytest -k filter]
TAG=${1:-san}
FILTER=${2:-"tc_ or attention or gru"}
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 0 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$FILTER" > gpurun_out/${TAG}_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Error" gpurun_out/${TAG}_memcheck.log | head -20
