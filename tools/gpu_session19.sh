#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session19.log) 2>&1
echo "=== probe (cluster=2)"; timeout 600 python tools/probe_conv.py v2
echo "=== bench_conv cluster=2"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v26_c2.json
echo "=== bench_conv cluster=1"; TTSB_CLUSTER=1 timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v26_c1.json
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256 cluster=2"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "=== bench b256 cluster=1"; TTSB_CLUSTER=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "=== done"
