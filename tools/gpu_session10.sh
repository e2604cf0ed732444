#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session11.log) 2>&1
echo "=== probe"; timeout 900 python tools/probe_conv.py v2
for L in s3_32_k3_d1 s2_64_k11_d5 s1_128_k11_d5 s0_256_k11_d5; do timeout 120 python tools/timeline.py $L; done
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --only s
echo "=== bench b256"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== launch list b64"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 200 --csv --log-file gpurun_out/launches_b64.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench64.log 2>&1
tail -1 gpurun_out/ncu_bench64.log | cut -c1-150
echo "=== done"
