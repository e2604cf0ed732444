#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session3.log) 2>&1
for L in s1_128_k11_d5 s1_128_k3_d1 s0_256_k3_d1 s3_32_k3_d1; do timeout 120 python tools/timeline.py $L; done
echo "=== per-tap mode (desc 3) for comparison"
TTSB_DESC_MODE=3 timeout 200 python tools/bench_conv.py --only s1
TTSB_DESC_MODE=3 timeout 120 python tools/timeline.py s1_128_k11_d5
echo "=== done"
