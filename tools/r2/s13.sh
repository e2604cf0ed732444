#!/bin/bash
# round 2, GPU session 13: conv_pair with two items per W2 pass (C = 64, k = 11)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s13.log) 2>&1
echo "=== pytest gpu (pair + models + variants + configs)"; timeout 1200 python -m pytest tests/test_gpu_convpair.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
echo "=== probe_pair x2"; timeout 300 python tools/probe_pair.py 2>&1 | grep "c64_k11\|c64_k7"
echo "=== probe_pair x1"; TTSB_PAIR_W2X2=0 timeout 300 python tools/probe_pair.py 2>&1 | grep "c64_k11\|c64_k7"
echo "=== timeline pair 64 11 5"; timeout 300 python tools/timeline_pair.py 64 11 5 | head -12
for i in 1 2; do
echo "=== vocoder alone, B=64, x2"; timeout 300 python tools/run_vocoder.py --batch 64 --reps 6 | cut -c1-160
echo "=== vocoder alone, B=64, x1"; TTSB_PAIR_W2X2=0 timeout 300 python tools/run_vocoder.py --batch 64 --reps 6 | cut -c1-160
done
echo "=== done"
