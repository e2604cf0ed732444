#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session5.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2
echo "=== bench_conv v2.1 (resident + rpp2)"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v21.json
echo "=== rpp=1"; TTSB_RPP=1 timeout 300 python tools/bench_conv.py --only s
echo "=== rpp=4"; TTSB_RPP=4 timeout 300 python tools/bench_conv.py --only s1
echo "=== resident off"; TTSB_RESIDENT=0 timeout 300 python tools/bench_conv.py --only s
echo "=== occ2=1"; TTSB_OCC2=1 timeout 300 python tools/bench_conv.py --only s
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
echo "=== done"
