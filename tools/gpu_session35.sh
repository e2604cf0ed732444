#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session35.log) 2>&1
echo "=== ncu source-level: conv_pair c64 k3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/src_pair_c64k3 \
   python tools/timeline_pair.py 64 3 1 8 > gpurun_out/ncu_src2.log 2>&1; tail -2 gpurun_out/ncu_src2.log
ls -la gpurun_out/*.ncu-rep
echo "=== done"
