#!/bin/bash
# round 2, second 2-GPU session: bench lines at N = 2 (target weak scaling, config 5 sharded) + the NCCL bit-identity test on the final kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s30.log) 2>&1
echo "=== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
echo "=== bench c5 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_s30_bench_c5_n2.json; cut -c1-300 gpurun_out/r2_s30_bench_c5_n2.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2_s30_bench_c5_n2.json
echo "=== bench c5 N=1"; timeout 900 python bench.py --gpus 1 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_s30_bench_c5_n1.json; cut -c1-300 gpurun_out/r2_s30_bench_c5_n1.json
echo "=== bench target N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2_s30_bench_target_n2.json; cut -c1-300 gpurun_out/r2_s30_bench_target_n2.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2_s30_bench_target_n2.json
echo "=== reference arm under torchrun N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-200
echo "=== done"
