#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session16.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2 simt
for L in s3_32_k3_d1 s1_128_k11_d5; do timeout 120 python tools/timeline.py $L | grep -v "tile [3-6]" | head -12; done
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v24.json
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== done"
