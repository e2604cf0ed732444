#!/bin/bash
# ncu --set full over one B=16 step with the end-of-round kernels: per-kernel DRAM traffic for roofline.traffic
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session52.log) 2>&1
timeout 1200 ncu --set full --clock-control none --import-source off -k regex:"conv_tc2|conv_pair|conv_post" -s 500 -c 125 \
    -o gpurun_out/r01_s52_full -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/r01_s52_full.ncu-rep --page raw --csv > gpurun_out/r01_s52_full_raw.csv 2>/dev/null
ls -la gpurun_out/r01_s52_full_raw.csv
rm -f gpurun_out/r01_s52_full.ncu-rep
echo "=== done"
