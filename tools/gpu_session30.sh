#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session30.log) 2>&1
nvidia-smi -L
echo "=== bench N=2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -3 > gpurun_out/bench_s30_n2.json; cut -c1-600 gpurun_out/bench_s30_n2.json
echo "=== reference arm N=2 (rank 0 only works)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-700
echo "=== bench N=1 full (with cpu baseline)"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s30_n1.json; cat gpurun_out/bench_s30_n1.json
echo "=== done"
