#!/bin/bash
# round 2, GPU session 6: persistent Tacotron2 decoder, device post-processing, CLI, e2e fix
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s6.log) 2>&1
echo "=== pytest tacotron2 + post + cli + api"; timeout 900 python -m pytest tests/test_gpu_tacotron2.py tests/test_gpu_t2_post.py tests/test_cli.py tests/test_gpu_api.py "tests/test_gpu_configs.py::test_config4_tacotron2_256_steps_vs_oracle" -x -q -m gpu 2>&1 | tail -15
echo "=== bench c4 (persistent)"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_s6_bench_c4.json; cut -c1-300 gpurun_out/r2_s6_bench_c4.json; grep -o '"decoder_us_per_step": [0-9.]*' gpurun_out/r2_s6_bench_c4.json
echo "=== bench c4 (six launches per step)"; TTSB_T2_PERSISTENT=0 timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*'
echo "=== bench target (e2e through parallel.synthesize, pinned)"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s6_bench_target.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2_s6_bench_target.json | head -6
echo "=== done"
