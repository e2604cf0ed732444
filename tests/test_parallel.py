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


def test_plan_shards_keeps_the_global_padding_condition():
    """Every shard whose longest utterance is shorter than the global maximum gets ONE extra padded token column, so
    that utterance still has a padded position to its right, as in the reference's single padded batch."""
    lengths = [9, 9, 7, 5, 3, 2]
    shards, pad_to = parallel.plan_shards(lengths, 2)
    assert sorted(i for s in shards for i in s) == list(range(6))
    assert pad_to == [9, 9]                      # both shards hold a globally longest utterance: no extra column
    shards, pad_to = parallel.plan_shards(lengths, 3)
    for idxs, p in zip(shards, pad_to):
        own = max(lengths[i] for i in idxs)
        assert p == (own if own == 9 else own + 1)
        for i in idxs:                           # padded iff padded in the global batch
            assert (lengths[i] < p) == (lengths[i] < max(lengths))
    assert parallel.plan_shards([4, 4], 4) == ([[0], [1], [], []], [4, 4, 0, 0])


class _FakeVocoder:
    hop = 4


class _FakeTTS:
    """Stands in for FastPitch2Wave on the CPU: utterance with ids x -> len(x)*2 frames -> waveform of value sum(x);
    records the padding arguments parallel.synthesize hands to the model."""
    vocoder = _FakeVocoder()
    device = torch.device('cpu')

    def __init__(self):
        self.seen = {}

    def synthesize_ids(self, id_list, speed, speaker_id, denoise, pitch_transform, max_duration, to_cpu=False, pad_to=0,
                       frame_len_hook=None, return_padded=False):
        from tts_arabic_pytorch_b200.models.fastpitch.networks import text_collate_fn
        padded, lens_sorted, inverse = text_collate_fn(id_list)
        t_local = int(lens_sorted[0]) * 2
        t = frame_len_hook(t_local)
        self.seen = {'pad_to': pad_to, 't_local': t_local, 't': t, 'l_sub': int(lens_sorted[0])}
        wav = torch.zeros(len(id_list), t * self.vocoder.hop)
        for r in range(len(id_list)):
            n = int(lens_sorted[r]) * 2 * self.vocoder.hop
            wav[r, :n] = float(padded[r].sum())
        return wav, lens_sorted * 2 * self.vocoder.hop, inverse, None


def _synth_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        lengths = [7, 3, 5, 2, 6]
        ids = [torch.randint(1, 40, (n,), generator=g) for n in lengths]
        model = _FakeTTS()
        res, stats = parallel.synthesize(model, ids, deliver='nccl', return_stats=True)
        # rank 0's shard holds the global longest utterance (7): no extra column / frame; rank 1's longest is 6
        want_pad = {0: (7, 14), 1: (7, 13)}[rank]
        ok = model.seen['pad_to'] == want_pad[0] and model.seen['t'] == want_pad[1] and model.seen['t_local'] in (14, 12)
        if rank == 0:
            ok = ok and len(res) == 5 and all(res[i].numel() == lengths[i] * 8 and
                                              bool((res[i] == float(ids[i].sum())).all()) for i in range(5))
        else:
            ok = ok and res is None
        q.put((rank, ok, stats['utterances']))
    finally:
        dist.destroy_process_group()


def test_synthesize_shards_pads_and_delivers_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_synth_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True, 3), (1, True, 2)]
