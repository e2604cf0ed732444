#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session4.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py simt v2 v1
eval "$(python tools/pick_mode.py)"
echo "=== chosen: impl=${TTSB_CONV_IMPL:-} version=${TTSB_TC_VERSION:-}"
echo "=== bench_conv v2"; TTSB_TC_VERSION=2 timeout 300 python tools/bench_conv.py --json gpurun_out/conv_v2.json
echo "=== bench_conv v1 (new epilogue)"; TTSB_TC_VERSION=1 timeout 300 python tools/bench_conv.py --only s
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
echo "=== done"
