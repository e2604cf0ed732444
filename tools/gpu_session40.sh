#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session40.log) 2>&1
echo "=== probe conv v2"; timeout 300 python tools/probe_conv.py v2
echo "=== bench_conv b32"; timeout 300 python tools/bench_conv.py --batch 32 --iters 7 --json gpurun_out/conv_s40_b32.json
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s40.json; cut -c1-200 gpurun_out/bench_s40.json
echo "=== done"
