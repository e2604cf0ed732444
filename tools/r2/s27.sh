#!/bin/bash
# round 2, GPU session 27: conv_pair C = 32 with two taps per K = 64 group in conv2 (TT rows [t | t+1], 128-byte swizzle)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s27.log) 2>&1
echo "=== pytest convpair"; timeout 900 python -m pytest tests/test_gpu_convpair.py -x -q -m gpu 2>&1 | tail -5
echo "=== probe_pair tt2"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep "c32"
echo "=== probe_pair TT2=0"; TTSB_PAIR_TT2=0 timeout 400 python tools/probe_pair.py --bench 2>&1 | grep "c32"
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s27_bench_target.json; cut -c1-300 gpurun_out/r2_s27_bench_target.json
echo "=== bench target TT2=0"; TTSB_PAIR_TT2=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-300
echo "=== done"
