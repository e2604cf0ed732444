#!/bin/bash
# round 2, GPU session 24: act-only epilogue on 32-column tiles (ups3); the full default bench line and the reference arm as the driver runs them
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s24.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench_conv ups"; timeout 300 python tools/bench_conv.py --batch 32 --only ups
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-900
echo "=== default bench (all legs)"; /usr/bin/time -v timeout 1200 python bench.py 2>gpurun_out/r2_s24_bench_stderr.txt | tail -1 > gpurun_out/r2_s24_bench_default.json; cut -c1-400 gpurun_out/r2_s24_bench_default.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/r2_s24_bench_default.json; grep -o '"gpu_eager_baseline": {[^}]*}' gpurun_out/r2_s24_bench_default.json; grep "Elapsed" gpurun_out/r2_s24_bench_stderr.txt; tail -3 gpurun_out/r2_s24_bench_stderr.txt
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== done"
