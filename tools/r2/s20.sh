#!/bin/bash
# round 2, GPU session 20: bisect the 4 failing tests of session 19 over the epilogue switches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s20.log) 2>&1
T="tests/test_gpu_configs.py::test_config3_b32_mel_and_waveforms_vs_oracle tests/test_gpu_models.py::test_full_size_round_trip_properties tests/test_gpu_models.py::test_mixed_length_batch_config5_shape"
echo "=== default"; timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed|Error" | head -30
echo "=== TMA_IN=0"; TTSB_EPI_TMA_IN=0 timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed" | head -12
echo "=== ACT=0"; TTSB_EPI_ACT=0 timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed" | head -12
echo "=== ACT=0 TMA_IN=0"; TTSB_EPI_ACT=0 TTSB_EPI_TMA_IN=0 timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed" | head -12
echo "=== PAIR_ACT_ONLY=0"; TTSB_PAIR_ACT_ONLY=0 timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed" | head -12
echo "=== RING=1"; TTSB_EPI_RING=1 timeout 600 python -m pytest $T -q -m gpu 2>&1 | grep -E "^E |passed|failed" | head -12
echo "=== done"
