#!/bin/bash
# round 2, GPU session 3: where the pair mid/final epilogues and the conv_tc2 issuer/epilogue spend their cycles
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s3.log) 2>&1
echo "=== fastpitch api tests (device-side id validation)"; timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -5
for cfg in "64 3 1" "32 3 1" "64 7 3" "64 11 5" "32 11 5"; do
  echo "=== timeline pair $cfg"; timeout 120 python tools/timeline_pair.py $cfg 16 | sed -n 1,11p
done
for l in s1_128_k11_d5 s1_128_k3_d1 s1_128_k7_d3 s0_256_k11_d5 s0_256_k3_d1; do
  echo "=== timeline tc2 $l"; timeout 120 python tools/timeline.py $l 32 2>/dev/null | sed -n 1,16p
done
echo "=== done"
