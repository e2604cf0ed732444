#!/bin/bash
# round 2, GPU session 31: conv_pair k = 3 timelines on the final kernels (what bounds the 1 900 / 2 500-cycle items?)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s31.log) 2>&1
for a in "32 3 1" "64 3 1" "64 7 3" "32 7 3"; do
  echo "=== timeline pair $a"; timeout 300 python tools/timeline_pair.py $a 2>/dev/null | head -11
done
echo "=== done"
