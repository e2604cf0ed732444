#!/bin/bash
# round 2, GPU session 7: tcgen05 attention, staged persistent Tacotron2 decoder
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s7.log) 2>&1
echo "=== pytest variants + models + configs + tacotron2"; timeout 1500 python -m pytest tests/test_gpu_variants.py tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_tacotron2.py tests/test_gpu_parallel.py -x -q -m gpu 2>&1 | tail -12
echo "=== bench c4 (persistent, staged)"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_s7_bench_c4.json; grep -o '"value": [0-9.]*\|"decoder_us_per_step": [0-9.]*' gpurun_out/r2_s7_bench_c4.json | head -3
echo "=== bench target: attention tc vs mma (per-kernel table)"
for att in tc mma; do
  TTSB_ATTENTION=$att timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s7_bench_target_$att.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_s7_bench_target_$att.json').read())
pk=d['roofline']['per_kernel']
print('$att', 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in pk.items() if k.startswith('fp.')})
PY
done
echo "=== done"
