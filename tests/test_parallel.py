"""N>1 host logic on CPU: world_size-2 gloo processes exercise sharding + waveform gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tts_arabic_pytorch_b200 import parallel


def test_snake_sharding_is_balanced_and_complete():
    lengths = [64 + (i * 37) % 193 for i in range(512)]
    shards = parallel.shard_utterances(lengths, 8)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(512))
    sums = [sum(lengths[i] for i in s) for s in shards]
    assert max(sums) - min(sums) <= max(lengths)
    assert all(len(s) == 64 for s in shards)
    # ragged: more ranks than utterances
    assert parallel.shard_utterances([5, 3], 4) == [[0], [1], [], []]
    items = [[lengths[i] for i in s] for s in shards]
    assert parallel.unshard(items, shards) == lengths


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lengths = [7, 3, 5, 2, 6]
        shards = parallel.shard_utterances(lengths, world)
        mine = shards[rank]
        n_max = max(lengths[i] for i in mine) * 4
        wav = torch.zeros(len(mine), n_max)
        for r, i in enumerate(mine):
            wav[r, :lengths[i] * 4] = float(i + 1)
        counts = torch.tensor([lengths[i] * 4 for i in mine])
        res = parallel.gather_waveforms(wav, counts, dst=0)
        if rank == 0:
            ws, cs = res
            per_rank = [[ws[r][j, :int(cs[r][j])] for j in range(ws[r].shape[0])] for r in range(world)]
            full = parallel.unshard(per_rank, shards)
            ok = all(full[i].numel() == lengths[i] * 4 and bool((full[i] == i + 1).all()) for i in range(len(lengths)))
            q.put(ok)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_gather_waveforms_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
