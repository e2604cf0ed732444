#!/bin/bash
# First GPU session: kernel-variant probe -> pick variant -> parity tests -> smoke -> bench -> launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session1.log) 2>&1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv
echo "=== probe"; timeout 900 python tools/probe_conv.py
eval "$(python tools/pick_mode.py)"
echo "=== chosen: impl=${TTSB_CONV_IMPL:-} desc_mode=${TTSB_DESC_MODE:-}"
echo "=== pytest gpu (continue past failures)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench b32"; timeout 600 python bench.py --steps 3 --warmup 3 --batch 32 --no-cpu-baseline 2>&1 | tail -3
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3
echo "=== ncu launch list (batch 4)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b4.csv \
    python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "=== done"
