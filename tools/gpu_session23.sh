#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session23.log) 2>&1
for only in s1_128 s0_256; do
echo "=== default $only b32"; timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
echo "=== cluster1 $only b32"; TTSB_CLUSTER=1 timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
echo "=== occ1 rpp2 $only"; TTSB_OCC2=1 TTSB_RPP=2 timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
echo "=== occ1 rpp2 cluster1 $only"; TTSB_CLUSTER=1 TTSB_OCC2=1 TTSB_RPP=2 timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
echo "=== occ1 rpp4 $only"; TTSB_OCC2=1 TTSB_RPP=4 timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
echo "=== occ1 rpp1 $only"; TTSB_OCC2=1 TTSB_RPP=1 timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
done
echo "=== done"
