#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session25.log) 2>&1
for dbg in 0 1 2 4 7 8 15; do
echo "=== TTSB_PAIR_DEBUG=$dbg"
TTSB_PAIR_DEBUG=$dbg timeout 120 python tools/timeline_pair.py 64 3 1 16 | sed -n 1,10p
TTSB_PAIR_DEBUG=$dbg timeout 120 python tools/timeline_pair.py 32 11 5 16 | sed -n 3,10p
done
echo "=== done"
