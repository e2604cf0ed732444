#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session33.log) 2>&1
echo "=== probe pair2"; TTSB_PAIR2=1 timeout 300 python tools/probe_conv.py v2
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
echo "=== bench_conv pair2 b32"; TTSB_PAIR2=1 timeout 300 python tools/bench_conv.py --batch 32 --iters 7
echo "=== bench_conv pair2 rpp1 b32"; TTSB_PAIR2=1 TTSB_PAIR2_RPP=1 timeout 300 python tools/bench_conv.py --batch 32 --iters 7 --only s1_128
echo "=== done"
