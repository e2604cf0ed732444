#!/bin/bash
# round 2: compute-sanitizer memcheck over the smoke path (FastPitch + generator, all round-2 epilogue kinds) and a Tacotron2 decode
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s35.log) 2>&1
echo "=== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8; echo "rc=$?"
echo "=== memcheck tacotron2 test"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tacotron2.py -x -q -m gpu -k "golden or oracle" 2>&1 | tail -6
echo "=== done"
