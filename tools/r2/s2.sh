#!/bin/bash
# round 2, GPU session 2: activated chain (one stored tensor per ResBlock input), pipelined mid epilogue, second staging tiles
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s2.log) 2>&1
echo "=== pytest gpu (pair + models + variants)"; timeout 900 python -m pytest tests/test_gpu_convpair.py tests/test_gpu_models.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -5
echo "=== probe_pair (activated form)"; timeout 600 python tools/probe_pair.py --bench --batch 16
echo "=== timeline pair c64 k3"; timeout 120 python tools/timeline_pair.py 64 3 1 16 | sed -n 1,10p
echo "=== timeline pair c32 k3"; timeout 120 python tools/timeline_pair.py 32 3 1 16 | sed -n 1,10p
echo "=== timeline pair c32 k11"; timeout 120 python tools/timeline_pair.py 32 11 5 16 | sed -n 1,10p
echo "=== vocoder alone, B=64"
timeout 300 python tools/run_vocoder.py --batch 64 --reps 5
TTSB_ACT_CHAIN=0 timeout 300 python tools/run_vocoder.py --batch 64 --reps 5
TTSB_STAGE2=0 timeout 300 python tools/run_vocoder.py --batch 64 --reps 5
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 126 -c 63 --csv --log-file gpurun_out/r2_s2_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 > /dev/null 2>&1
grep -c conv gpurun_out/r2_s2_launches_b64.csv
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_s2_bench.json; cut -c1-300 gpurun_out/r2_s2_bench.json
echo "=== done"
