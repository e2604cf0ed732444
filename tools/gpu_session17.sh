#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session17.log) 2>&1
echo "=== occ1 rpp2"; TTSB_OCC2=1 TTSB_RPP=2 timeout 300 python tools/bench_conv.py --only s
echo "=== occ1 rpp1"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --only s
echo "=== rpp2 (default occ)"; TTSB_RPP=2 timeout 300 python tools/bench_conv.py --only s
echo "=== rpp4 (default occ)"; TTSB_RPP=4 timeout 300 python tools/bench_conv.py --only s
echo "=== occ3"; TTSB_OCC2=3 timeout 300 python tools/bench_conv.py --only s
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1
echo "=== done"
