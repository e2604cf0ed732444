#!/bin/bash
# round 2, GPU session 19: TMA-in refill before the math (with the missing warp reconvergence), act-only for rpp > 1 and raw-only launches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s19.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "=== pytest models x3 (race check)"; for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu 2>&1 | tail -1; done
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --batch 32 | grep -v "s2_\|s3_"
echo "=== timeline tc2 s1_128_k3_d1"; timeout 300 python tools/timeline.py s1_128_k3_d1 16 2>/dev/null | head -12
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s19_bench_target.json; cut -c1-300 gpurun_out/r2_s19_bench_target.json
echo "=== done"
