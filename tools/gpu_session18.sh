#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session18.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2
echo "=== bench_conv"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v25.json
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "=== launch list b64 (one chunk)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/launches_b64.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench64.log 2>&1
tail -1 gpurun_out/ncu_bench64.log | cut -c1-120
echo "=== done"
