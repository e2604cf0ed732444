#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session42.log) 2>&1
echo "=== probe_pair"; timeout 600 python tools/probe_pair.py --bench --batch 16 | grep -v "_d[35] "
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s42.json; cut -c1-200 gpurun_out/bench_s42.json
echo "=== launch list b64"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_pair -s 54 -c 18 --csv --log-file gpurun_out/launches_s42_pair.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench64.log 2>&1
grep conv_pair gpurun_out/launches_s42_pair.csv | awk -F'","' '{print $5, $NF}' | cut -c1-30,90-120
echo "=== done"
