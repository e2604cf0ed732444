#!/bin/bash
# round 2, GPU session 5: new parity tests at config sizes, sharded == single batch, new bench.py (all configs)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s5.log) 2>&1
echo "=== pytest gpu (new files first)"; timeout 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -15
echo "=== pytest gpu (everything)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in target c2 c3 c4 c5; do
  echo "=== bench $cfg"; timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 2>&1 | tail -2 > gpurun_out/r2_s5_bench_$cfg.json; cut -c1-600 gpurun_out/r2_s5_bench_$cfg.json
done
echo "=== done"
