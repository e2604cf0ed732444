#!/bin/bash
# round 2, GPU session 14: act-only lean epilogue (kEpi 3 / 4) A/B
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s14.log) 2>&1
echo "=== pytest gpu (pair + models + variants + configs), EPI_ACT=2"; TTSB_EPI_ACT=2 timeout 1200 python -m pytest tests/test_gpu_convpair.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
echo "=== pytest gpu (all), default"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for m in 0 1 2; do
  echo "=== bench_conv TTSB_EPI_ACT=$m"; TTSB_EPI_ACT=$m timeout 300 python tools/bench_conv.py --batch 32 --only _ | grep -v "s2_\|s3_\|ff\|qkv"
done
for l in s1_128_k3_nores; do
  echo "=== timeline tc2 $l"; timeout 300 python tools/timeline.py $l 16 2>/dev/null | head -12
done
echo "=== timeline tc2 s1_128_k3_d1 EPI_ACT=2"; TTSB_EPI_ACT=2 timeout 300 python tools/timeline.py s1_128_k3_d1 16 2>/dev/null | head -12
for i in 1 2 3; do
for m in 0 1 2; do
echo "=== vocoder alone, B=64, EPI_ACT=$m"; TTSB_EPI_ACT=$m timeout 300 python tools/run_vocoder.py --batch 64 --reps 4 | cut -c1-160
done
done
echo "=== done"
