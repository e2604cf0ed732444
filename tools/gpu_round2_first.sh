#!/bin/bash
# First GPU call of the next round ("What bounds the C >= 128 layers, and the next experiments"):
#   1. today's kernels under the occupancy / rows-per-pass knobs (last measured before the epilogue fixes, profiles/r01_s23)
#   2. the rotated weight-tile order / fast issue path, if tools/experiments/{fast_issue,rotate_taps}_conv_tc2.patch were applied and the
#      library rebuilt HERE before the call (TTSB_ROTATE_TAPS is ignored by an unpatched build)
# ~3 GPU-minutes. Output: gpurun_out/round2_first.log
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/round2_first.log) 2>&1
run() { echo "=== $1"; shift; env "$@" timeout 300 python tools/bench_conv.py --batch 32 --only "$ONLY" --iters 7; }
for ONLY in s1_128 s0_256; do
    run "default $ONLY" TTSB_NOP=1
    run "rotate $ONLY" TTSB_ROTATE_TAPS=1
    run "occ1 rpp2 $ONLY" TTSB_OCC2=1 TTSB_RPP=2
    run "occ1 rpp2 rotate $ONLY" TTSB_OCC2=1 TTSB_RPP=2 TTSB_ROTATE_TAPS=1
    run "occ1 rpp1 $ONLY" TTSB_OCC2=1 TTSB_RPP=1
done
# mma[gotTMEM gotA issued]: gotA - gotTMEM = wait for the A panel; (issued - gotA) / (n_chunks x n_taps) = weight-ring period
echo "=== in-kernel timelines of the k=11 layers"
timeout 120 python tools/timeline.py s1_128_k11_d5 32 | head -40
timeout 120 python tools/timeline.py s0_256_k11_d5 32 | head -40
# conv_pair: k >= 7 items take ~2 x the summed MMA time although its issue loops are lean; find the stage that is late
echo "=== conv_pair timelines, k = 11 / 7"
timeout 120 python tools/timeline_pair.py 64 11 5 | head -30
timeout 120 python tools/timeline_pair.py 64 7 3 | head -30
timeout 120 python tools/timeline_pair.py 32 11 5 | head -30
echo "=== vocoder parity with the rotated order (golden + oracle + SIMT cross-check)"
TTSB_ROTATE_TAPS=1 timeout 600 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "hifigan or tcgen05" 2>&1 | tail -3
echo "=== done"
