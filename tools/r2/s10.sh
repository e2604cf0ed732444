#!/bin/bash
# round 2, GPU session 10: same-box A/B of the hoisted residual pointers (TTSB_LIB), persistent-decoder phase breakdown
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s10.log) 2>&1
BASE=$PWD/tts_arabic_pytorch_b200/libttsb200_base.so
for i in 1 2 3; do
  echo "--- base"; TTSB_LIB=$BASE timeout 300 python tools/run_vocoder.py --batch 64 --reps 6 | cut -c1-160
  echo "--- new";  timeout 300 python tools/run_vocoder.py --batch 64 --reps 6 | cut -c1-160
done
echo "=== t2 phases"; timeout 300 python tools/t2_phases.py 8 256
timeout 300 python tools/t2_phases.py 1 128
echo "=== done"
