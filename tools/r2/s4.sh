#!/bin/bash
# round 2, GPU session 4: shared-memory residual in conv_pair, scoreboard-order fix + early bias in the lean epilogue
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s4.log) 2>&1
echo "=== pytest gpu (pair + models + variants + api)"; timeout 900 python -m pytest tests/test_gpu_convpair.py tests/test_gpu_models.py tests/test_gpu_variants.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -5
echo "=== probe_pair (activated form)"; timeout 600 python tools/probe_pair.py --bench --batch 16 | grep -v "^device"
for cfg in "64 3 1" "32 3 1" "32 11 5"; do
  echo "=== timeline pair $cfg"; timeout 120 python tools/timeline_pair.py $cfg 16 | sed -n 1,11p
done
for l in s1_128_k11_d5 s1_128_k3_d1 s0_256_k7_d3; do
  echo "=== timeline tc2 $l"; timeout 120 python tools/timeline.py $l 32 2>/dev/null | sed -n 1,12p
done
echo "=== bench_conv"; timeout 300 python tools/bench_conv.py --batch 32 --iters 7
echo "=== vocoder alone, B=64"
timeout 300 python tools/run_vocoder.py --batch 64 --reps 6
echo "=== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s4_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > /dev/null 2>&1
grep -c conv gpurun_out/r2_s4_launches_b64.csv
echo "=== bench b256"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_s4_bench.json; cut -c1-300 gpurun_out/r2_s4_bench.json
echo "=== done"
