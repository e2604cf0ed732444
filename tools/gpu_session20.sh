#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session20.log) 2>&1
echo "=== probe_pair"; timeout 600 python tools/probe_pair.py --bench --batch 16
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "=== bench b256 pair"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "=== bench b256 nopair"; TTSB_PAIR=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "=== done"
