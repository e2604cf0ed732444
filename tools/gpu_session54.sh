#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session54.log) 2>&1
echo "=== pytest"; timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "=== bench b256"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s54.json; cut -c1-200 gpurun_out/bench_s54.json
echo "=== done"
