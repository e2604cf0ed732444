#!/bin/bash
# round 2, GPU session 16: TMA-in epilogue (kEpi 4..7) A/B
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s16.log) 2>&1
echo "=== pytest gpu (variants first)"; timeout 1200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_convpair.py -x -q -m gpu 2>&1 | tail -5
echo "=== plans"; TTSB_EPI_VERBOSE=1 timeout 300 python tools/run_vocoder.py --batch 8 --reps 1 2>&1 | grep "conv_tc2:" | sort | uniq -c
for m in 0 1; do
  echo "=== bench_conv TTSB_EPI_TMA_IN=$m"; TTSB_EPI_TMA_IN=$m timeout 300 python tools/bench_conv.py --batch 32 --only _k | grep -v "s2_\|s3_\|ff\|qkv"
done
echo "=== bench_conv TTSB_EPI_RING=1"; TTSB_EPI_RING=1 timeout 300 python tools/bench_conv.py --batch 32 --only _k | grep -v "s2_\|s3_\|ff\|qkv"
echo "=== timeline tc2 s1_128_k3_d1"; timeout 300 python tools/timeline.py s1_128_k3_d1 16 2>/dev/null | head -12
echo "=== timeline tc2 s1_128_k7_d3"; timeout 300 python tools/timeline.py s1_128_k7_d3 16 2>/dev/null | head -12
for i in 1 2 3; do
for m in 0 1; do
echo "=== vocoder alone, B=64, TMA_IN=$m"; TTSB_EPI_TMA_IN=$m timeout 300 python tools/run_vocoder.py --batch 64 --reps 4 | cut -c1-160
done
done
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s16_bench_target.json; cut -c1-300 gpurun_out/r2_s16_bench_target.json
echo "=== done"
