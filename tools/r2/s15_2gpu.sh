#!/bin/bash
# round 2, 2-GPU session: the sharded product path over NCCL (bit-identity against the single padded batch), config 5 and
# the target config at N = 2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s15.log) 2>&1
nvidia-smi -L
echo "=== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
echo "=== check script output"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/check_parallel_nccl.py 2>&1 | grep -v "^W\|^\*\|OMP" | tail -5
echo "=== bench c5 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s15_bench_c5_n2.json; cut -c1-400 gpurun_out/r2_s15_bench_c5_n2.json
echo "=== bench c5 N=1"; timeout 900 python bench.py --gpus 1 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s15_bench_c5_n1.json; cut -c1-400 gpurun_out/r2_s15_bench_c5_n1.json
echo "=== bench target N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r2_s15_bench_target_n2.json; cut -c1-400 gpurun_out/r2_s15_bench_target_n2.json
echo "=== done"
