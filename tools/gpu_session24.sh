#!/bin/bash
# ncu --set full over one B=64 step (one vocoder chunk): per-kernel DRAM traffic, tensor-pipe and smem utilisation
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session24.log) 2>&1
echo "=== ncu full b64"
timeout 1500 ncu --set full --clock-control none --import-source off -k regex:"conv_tc2|conv_pair|conv_post" -s 560 -c 125 \
    -o gpurun_out/r01_s24_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
echo "=== raw csv"
ncu -i gpurun_out/r01_s24_full.ncu-rep --page raw --csv > gpurun_out/r01_s24_full_raw.csv 2>/dev/null
ls -la gpurun_out/r01_s24_full_raw.csv
# the report itself may exceed the return limit; keep it only if small
sz=$(stat -c %s gpurun_out/r01_s24_full.ncu-rep)
if [ "$sz" -gt 45000000 ]; then rm -f gpurun_out/r01_s24_full.ncu-rep; echo "report dropped ($sz bytes)"; fi
echo "=== done"
