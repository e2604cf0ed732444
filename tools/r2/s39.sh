#!/bin/bash
# round 2, GPU session 39: conv_pair with two conv2 issuers (even / odd items) when W2 is resident
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s39.log) 2>&1
echo "=== pytest convpair + models + variants + configs"; timeout 1200 python -m pytest tests/test_gpu_convpair.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
echo "=== probe_pair split"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== probe_pair one conv2 issuer"; TTSB_PAIR_C2_SPLIT=0 timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== timeline pair 32 3 1"; timeout 300 python tools/timeline_pair.py 32 3 1 2>/dev/null | head -11
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s39_bench_target.json; cut -c1-300 gpurun_out/r2_s39_bench_target.json
echo "=== bench target, one conv2 issuer"; TTSB_PAIR_C2_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-300
echo "=== done"
