#!/bin/bash
# round 2, GPU session 21: TMA-in early refill behind a proxy fence; overlapped D2H in the API path
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s21.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "=== pytest models+variants x2 (race check)"; for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_variants.py -q -m gpu 2>&1 | tail -1; done
echo "=== bench_conv"; timeout 300 python tools/bench_conv.py --batch 32 --only s1_128
echo "=== bench_conv OCC2=1 s0"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --batch 32 --only s0_256
echo "=== bench_conv default s0"; timeout 300 python tools/bench_conv.py --batch 32 --only s0_256
echo "=== timeline tc2 s1_128_k3_d1"; timeout 300 python tools/timeline.py s1_128_k3_d1 16 2>/dev/null | head -12
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s21_bench_target.json; cut -c1-300 gpurun_out/r2_s21_bench_target.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2_s21_bench_target.json
echo "=== done"
