#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session46.log) 2>&1
echo "=== probe conv v2 (tma out)"; timeout 300 python tools/probe_conv.py v2
echo "=== bench_conv b32 tma out"; timeout 300 python tools/bench_conv.py --batch 32 --iters 7
echo "=== bench_conv b32 no tma out"; TTSB_TMA_OUT=0 timeout 300 python tools/bench_conv.py --batch 32 --iters 7 --only _k3
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s46.json; cut -c1-200 gpurun_out/bench_s46.json
echo "=== done"
