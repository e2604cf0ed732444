#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session15.log) 2>&1
echo "=== rpp2"; TTSB_RPP=2 timeout 300 python tools/bench_conv.py --only s
echo "=== rpp4"; TTSB_RPP=4 timeout 300 python tools/bench_conv.py --only s
echo "=== probe rpp4"; TTSB_RPP=4 timeout 600 python tools/probe_conv.py v2
echo "=== pytest (incl. tacotron2)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "=== done"
