#!/bin/bash
# Session 2: per-layer timing under tuning knobs + ncu full captures of three representative launches.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session2.log) 2>&1
echo "=== baseline"; timeout 300 python tools/bench_conv.py --json gpurun_out/conv_base.json
echo "=== smem budget 110000 (2 CTAs/SM where possible)"; TTSB_SMEM_BUDGET=110000 timeout 300 python tools/bench_conv.py
echo "=== smem budget 70000"; TTSB_SMEM_BUDGET=70000 timeout 300 python tools/bench_conv.py --only s
echo "=== max b stages 2"; TTSB_MAX_B_STAGES=2 timeout 300 python tools/bench_conv.py --only s
echo "=== debug: no epilogue traffic"; TTSB_DEBUG_FLAGS=1 timeout 300 python tools/bench_conv.py --only s
echo "=== debug: no weight streaming"; TTSB_DEBUG_FLAGS=2 timeout 300 python tools/bench_conv.py --only s
echo "=== debug: neither"; TTSB_DEBUG_FLAGS=3 timeout 300 python tools/bench_conv.py --only s
for L in s1_128_k11_d5 s3_32_k3_d1 s0_256_k3_d1; do
  echo "=== ncu $L"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 \
      -o gpurun_out/prof_$L -f python tools/bench_conv.py --only $L --iters 2 > gpurun_out/ncu_$L.log 2>&1
  tail -2 gpurun_out/ncu_$L.log
done
ls -la gpurun_out
echo "=== done"
