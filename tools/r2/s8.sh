#!/bin/bash
# round 2, GPU session 8: optimized persistent decoder, ragged generator chunks (config 5)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s8.log) 2>&1
echo "=== pytest (tacotron2, configs, models, api, parallel)"; timeout 1500 python -m pytest tests/test_gpu_tacotron2.py tests/test_gpu_configs.py tests/test_gpu_models.py tests/test_gpu_api.py tests/test_gpu_parallel.py tests/test_gpu_denoiser.py -x -q -m gpu 2>&1 | tail -6
echo "=== bench c4"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_s8_bench_c4.json; grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*' gpurun_out/r2_s8_bench_c4.json | head -3
echo "=== bench c5 (N=1, ragged chunks)"; timeout 900 python bench.py --config c5 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_s8_bench_c5.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2_s8_bench_c5.json | head -4
echo "=== bench target"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s8_bench_target.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2_s8_bench_target.json | head -4
echo "=== done"
