#!/bin/bash
# round 2, GPU session 25: the default bench line (all legs) as the driver runs it
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s25.log) 2>&1
echo "=== default bench (all legs)"; S=$(date +%s); timeout 1200 python bench.py 2>gpurun_out/r2_s25_bench_stderr.txt | tail -1 > gpurun_out/r2_s25_bench_default.json; echo "wall $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/r2_s25_bench_default.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/r2_s25_bench_default.json; grep -o '"gpu_eager_baseline": {[^}]*}' gpurun_out/r2_s25_bench_default.json; tail -3 gpurun_out/r2_s25_bench_stderr.txt
for c in c2 c3 c4 c5; do echo "=== bench $c"; timeout 900 python bench.py --config $c --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_s25_bench_$c.json; cut -c1-250 gpurun_out/r2_s25_bench_$c.json; done
echo "=== done"
