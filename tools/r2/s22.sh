#!/bin/bash
# round 2, GPU session 22: one vs two CTAs per SM across the layer table (which plans should prefer occupancy 1?)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s22.log) 2>&1
echo "=== default"; timeout 300 python tools/bench_conv.py --batch 32 | grep -v "s2_\|s3_"
echo "=== OCC2=1"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --batch 32 | grep -v "s2_\|s3_"
echo "=== done"
