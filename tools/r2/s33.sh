#!/bin/bash
# round 2, GPU session 33: persistent Tacotron2 decoder with the query / projection / prenet rows spread over all CTAs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s33.log) 2>&1
echo "=== pytest tacotron2 + config4 + api"; timeout 900 python -m pytest tests/test_gpu_tacotron2.py "tests/test_gpu_configs.py::test_config4_tacotron2_256_steps_vs_oracle" tests/test_gpu_t2_post.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -4
echo "=== t2 phases"; timeout 300 python tools/t2_phases.py 8 256
timeout 300 python tools/t2_phases.py 1 128
echo "=== bench c4"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_s33_bench_c4.json; grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*' gpurun_out/r2_s33_bench_c4.json | head -3
echo "=== done"
