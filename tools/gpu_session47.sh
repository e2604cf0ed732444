#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session47.log) 2>&1
echo "=== probe_pair"; timeout 600 python tools/probe_pair.py --bench --batch 16
echo "=== timeline"; timeout 120 python tools/timeline_pair.py 64 3 1 16 | sed -n 1,10p
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s47.json; cut -c1-200 gpurun_out/bench_s47.json
echo "=== done"
