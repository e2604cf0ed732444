#!/bin/bash
# round 2, GPU session 29: tt_pair as a template parameter (the run-time flag slowed the C = 64 instantiations)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s29.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== probe_pair"; timeout 400 python tools/probe_pair.py --bench 2>&1 | grep " us "
echo "=== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_s29_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > /dev/null 2>&1
grep -c "conv" gpurun_out/r02_s29_launches_b64.csv
echo "=== bench target"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/r2_s29_bench_target.json; cut -c1-300 gpurun_out/r2_s29_bench_target.json
echo "=== done"
