#!/bin/bash
# round 2, GPU session 1: sanity of the round-1 tree on a fresh box + MMA issue-loop variants (TTSB_ISSUE=0/1/2) + pair transform unroll
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s1.log) 2>&1
nvidia-smi -L
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for mode in 0 1 2; do
  for only in s1_128 s0_256 ups ff; do
    echo "=== issue mode $mode $only"
    TTSB_ISSUE=$mode timeout 300 python tools/bench_conv.py --batch 32 --only $only --iters 7
  done
done
echo "=== probe_pair"; timeout 600 python tools/probe_pair.py --bench --batch 16
echo "=== timeline pair c64 k3"; timeout 120 python tools/timeline_pair.py 64 3 1 16 | sed -n 1,10p
echo "=== timeline pair c32 k3"; timeout 120 python tools/timeline_pair.py 32 3 1 16 | sed -n 1,10p
echo "=== timeline pair c64 k11"; timeout 120 python tools/timeline_pair.py 64 11 5 16 | sed -n 1,10p
echo "=== timeline tc2 s1 k11 (mode 2)"; timeout 120 python tools/timeline.py s1_128_k11_d5 32 | head -30
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_s1_bench.json; cut -c1-400 gpurun_out/r2_s1_bench.json
echo "=== done"
