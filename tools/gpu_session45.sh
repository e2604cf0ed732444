#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session45.log) 2>&1
echo "=== probe conv v2"; timeout 300 python tools/probe_conv.py v2
echo "=== probe_pair"; timeout 600 python tools/probe_pair.py --bench --batch 16 | grep -v "_d[35] "
echo "=== bench_conv b32"; timeout 300 python tools/bench_conv.py --batch 32 --iters 7
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s45.json; cut -c1-200 gpurun_out/bench_s45.json
echo "=== done"
