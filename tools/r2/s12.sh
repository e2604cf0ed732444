#!/bin/bash
# round 2, GPU session 12: where does the lean epilogue's time go? steady-chunk stamps + timing-decomposition switches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s12.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for l in s1_128_k3_nores s1_128_k3_d1 s1_128_k7_nores; do
  echo "=== timeline tc2 $l"; timeout 300 python tools/timeline.py $l 16 | head -20
done
DBG=$PWD/tts_arabic_pytorch_b200/libttsb200_dbg.so
echo "=== production library"; timeout 300 python tools/bench_conv.py --batch 32 --only s1_128_k
for d in 0 1 2 4 8 3 15; do
  echo "=== debug library TTSB_EPI_DEBUG=$d"; TTSB_LIB=$DBG TTSB_EPI_DEBUG=$d timeout 300 python tools/bench_conv.py --batch 32 --only s1_128_k
done
echo "=== TMA_OUT=0"; TTSB_TMA_OUT=0 timeout 300 python tools/bench_conv.py --batch 32 --only s1_128_k
echo "=== OCC2=1"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --batch 32 --only s1_128_k
echo "=== done"
