#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session6.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v22.json
echo "=== occ2=1"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --only s
echo "=== occ1 rpp2"; TTSB_OCC2=1 TTSB_RPP=2 timeout 300 python tools/bench_conv.py --only s
echo "=== bench b256 chunk default"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== bench b256 chunk 8192"; TTSB_HIFIGAN_CHUNK_FRAMES=8192 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== bench b256 chunk 32768"; TTSB_HIFIGAN_CHUNK_FRAMES=32768 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== done"
