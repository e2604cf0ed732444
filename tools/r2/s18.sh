#!/bin/bash
# round 2, GPU session 18: TMA-in refill issued before the math; conv_pair act-only final epilogue A/B
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s18.log) 2>&1
echo "=== pytest gpu (variants, models, configs, pair)"; timeout 1200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_convpair.py -x -q -m gpu 2>&1 | tail -3
echo "=== bench_conv default"; timeout 300 python tools/bench_conv.py --batch 32 --only s1_128
echo "=== timeline tc2 s1_128_k3_d1"; timeout 300 python tools/timeline.py s1_128_k3_d1 16 2>/dev/null | head -12
echo "=== probe_pair act-only"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== probe_pair round-1 final epilogue"; TTSB_PAIR_ACT_ONLY=0 timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_s18_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > /dev/null 2>&1
grep -c "conv" gpurun_out/r02_s18_launches_b64.csv
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s18_bench_target.json; cut -c1-300 gpurun_out/r2_s18_bench_target.json
echo "=== bench target, pair act-only off"; TTSB_PAIR_ACT_ONLY=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel 2>/dev/null | tail -1 | cut -c1-300
echo "=== done"
