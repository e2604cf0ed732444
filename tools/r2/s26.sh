#!/bin/bash
# round 2, GPU session 26: LayerNorm / non-LayerNorm halves of the general epilogue; final numbers of the single-GPU configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s26.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s26_bench_target.json; cut -c1-300 gpurun_out/r2_s26_bench_target.json
echo "=== default bench (all legs)"; S=$(date +%s); timeout 1200 python bench.py 2>gpurun_out/r2_s26_bench_stderr.txt | tail -1 > gpurun_out/r2_s26_bench_default.json; echo "wall $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/r2_s26_bench_default.json; grep -o '"gpu_eager_baseline": {[^}]*}' gpurun_out/r2_s26_bench_default.json
echo "=== done"
