#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session13.log) 2>&1
for L in s3_32_k3_d1 s1_128_k11_d5; do timeout 120 python tools/timeline.py $L | grep -v "cta   1 \|cta 147\|cta 200\|cta 255" | head -60; done
echo "=== done"
