#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session43.log) 2>&1
echo "=== launch list b64 (one vocoder chunk per step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 170 --csv --log-file gpurun_out/launches_s43_b64.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench64.log 2>&1
tail -1 gpurun_out/ncu_bench64.log | cut -c1-120
echo "=== done"
