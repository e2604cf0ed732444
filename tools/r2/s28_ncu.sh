#!/bin/bash
# round 2, final profiler evidence: launch list + ncu --set full over ONE generator pass at batch 64 x 512 frames (final kernels)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s28.log) 2>&1
echo "=== launch list (gpu__time_duration), one pass"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_s28_launches_b64.csv python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > /dev/null 2>&1
grep -c "conv" gpurun_out/r02_s28_launches_b64.csv
echo "=== ncu --set full, one pass"
timeout 1500 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r02_s28_full_b64 -f python tools/run_vocoder.py --batch 64 --reps 3 --profile-last > gpurun_out/r02_s28_ncu_full.log 2>&1
ls -la gpurun_out/r02_s28_full_b64.ncu-rep
timeout 600 ncu -i gpurun_out/r02_s28_full_b64.ncu-rep --page raw --csv > gpurun_out/r02_s28_full_b64_raw.csv 2>/dev/null
wc -c gpurun_out/r02_s28_full_b64_raw.csv
rm -f gpurun_out/r02_s28_full_b64.ncu-rep
echo "=== whole-step launch list (B=64, FastPitch + generator), bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s28_step_launches_b64.csv python bench.py --config c3 --batch 64 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-per-kernel > /dev/null 2>&1
wc -l gpurun_out/r02_s28_step_launches_b64.csv
echo "=== done"
