#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session50.log) 2>&1
nvidia-smi -L
echo "=== pytest"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench N=1 full"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s50_n1.json; cat gpurun_out/bench_s50_n1.json
echo "=== bench N=2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s50_n2.json; cut -c1-300 gpurun_out/bench_s50_n2.json
echo "=== done"
