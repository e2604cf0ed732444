#!/bin/bash
# round 2, GPU session 23: occupancy-1 plan for the C = 256 single-tile layers; shared-memory residual with a 3-deep x ring (C = 64, k = 7)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s23.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== probe_pair default"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep "c64_k7\|c64_k11"
echo "=== probe_pair SMEM_RES_SLOTS=3"; TTSB_PAIR_SMEM_RES_SLOTS=3 timeout 400 python tools/probe_pair.py --bench 2>&1 | grep "c64_k7\|c64_k11"
echo "=== probe_pair SMEM_RES_SLOTS=2"; TTSB_PAIR_SMEM_RES_SLOTS=2 timeout 400 python tools/probe_pair.py --bench 2>&1 | grep "c64_k7\|c64_k11"
echo "=== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_s23_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > /dev/null 2>&1
grep -c "conv" gpurun_out/r02_s23_launches_b64.csv
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s23_bench_target.json; cut -c1-300 gpurun_out/r2_s23_bench_target.json
echo "=== bench target SMEM_RES_SLOTS=3"; TTSB_PAIR_SMEM_RES_SLOTS=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-300
echo "=== done"
