#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session9.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v23.json
echo "=== occ2=3"; TTSB_OCC2=3 timeout 300 python tools/bench_conv.py --only s
echo "=== probe occ3"; TTSB_OCC2=3 timeout 900 python tools/probe_conv.py v2
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== bench b256 occ3"; TTSB_OCC2=3 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== done"
