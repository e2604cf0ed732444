#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session32.log) 2>&1
echo "=== ncu source-level: conv_tc2 s1_128_k7 b8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 12 -c 1 -f -o gpurun_out/src_tc2_c128k7 \
   python tools/bench_conv.py --batch 8 --iters 12 --only s1_128_k7 > gpurun_out/ncu_src1.log 2>&1; tail -2 gpurun_out/ncu_src1.log
echo "=== ncu source-level: conv_pair c64 k3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 4 -c 1 -f -o gpurun_out/src_pair_c64k3 \
   python tools/timeline_pair.py 64 3 1 8 > gpurun_out/ncu_src2.log 2>&1; tail -2 gpurun_out/ncu_src2.log
ls -la gpurun_out/*.ncu-rep
echo "=== done"
