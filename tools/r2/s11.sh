#!/bin/bash
# round 2, GPU session 11: persistent decoder with batched weight loads; 2-GPU runs come next
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s11.log) 2>&1
echo "=== pytest tacotron2 + config4"; timeout 900 python -m pytest tests/test_gpu_tacotron2.py "tests/test_gpu_configs.py::test_config4_tacotron2_256_steps_vs_oracle" tests/test_gpu_t2_post.py -x -q -m gpu 2>&1 | tail -4
echo "=== t2 phases"; timeout 300 python tools/t2_phases.py 8 256
echo "=== bench c4"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_s11_bench_c4.json; grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*' gpurun_out/r2_s11_bench_c4.json | head -3
echo "=== bench c4 legacy six launches"; TTSB_T2_PERSISTENT=0 timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*' | head -2
echo "=== done"
