"""The N > 1 product path on real GPUs: one process per GPU over NCCL (tools/check_parallel_nccl.py under
torch.distributed.run). Needs at least two visible GPUs; a one-GPU box skips it (tests/test_gpu_parallel.py checks the
same sharding arithmetic rank by rank on one GPU, tests/test_parallel.py the host logic over gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_synthesis_over_nccl_is_bit_identical():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (they never fall back to the CPU)')
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(REPO, 'tools', 'check_parallel_nccl.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'parallel nccl check: PASS' in r.stdout, r.stdout[-3000:]
