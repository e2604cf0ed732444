#!/bin/bash
# round 2, GPU session 32: mbarrier waits with the default try_wait time limit instead of the 100 us suspend hint (variant library)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s32.log) 2>&1
SPIN=$PWD/tts_arabic_pytorch_b200/libttsb200_spin.so
echo "=== probe_pair default"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== probe_pair spin"; TTSB_LIB=$SPIN timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --batch 32 | grep -v "s2_\|s3_"
echo "=== bench_conv spin"; TTSB_LIB=$SPIN timeout 300 python tools/bench_conv.py --batch 32 | grep -v "s2_\|s3_"
echo "=== bench target default"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-200
echo "=== bench target spin"; TTSB_LIB=$SPIN timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-200
echo "=== done"
