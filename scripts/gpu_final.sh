#!/bin/bash
# Evidence run of a build on one B200: GPU tests, the bench line, the ncu launch list of one eager step, `--set full` captures of
# the top kernels, and the DRAM traffic of every launch of the dominant kernel.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
# launch list of a short eager run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full-set captures: one launch of each hot kernel from the last steps of the same command
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'tc_gemm_ts_kernel|tc_wgrad_ts_kernel|attn_fwd|attn_bwd|gru_scan' -s 560 -c 40 \
  -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
# the report itself is ~2.5 MB per kernel with sources (gpurun_out/ is capped at 64 MiB): keep the raw metrics page as CSV, and
# the report only when small
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/${TAG}_prof.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/${TAG}_prof.ncu-rep; fi
# DRAM traffic of every dense forward / backward-data launch
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'tc_gemm_ts_kernel' -c 600 --csv --log-file gpurun_out/${TAG}_traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --no-extra > gpurun_out/${TAG}_traffic.log 2>&1
# rollouts: launch lists of 6 timesteps of BASELINE config 2 (fused FF acting kernel) and config 4 (RNN agent, tensor-core path)
LIMIT=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_rollout_cfg2_launches.csv python scripts/prof_rollout.py refil_group_matching 4096 4 4 > /dev/null 2>&1
LIMIT=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_rollout_cfg4_launches.csv python scripts/prof_rollout.py refil 4096 8 24 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
tail -1 gpurun_out/${TAG}_smoke.log
ls -la gpurun_out/ | grep ${TAG}
