"""Multi-GPU layer: utterances are independent end to end, so a batch shards across ranks with no
bulk data-path collective; the only bulk exchange is the final delivery of waveforms.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).
  shard_utterances   length-balanced "snake" deal of a length-sorted utterance list
  plan_shards        + the padding each shard needs to keep the GLOBAL batch's padded-position condition
  synthesize         the product entry point: FastPitch2Wave over a sharded utterance list, results on `dst`
  gather_waveforms   all_gather of sample counts, then one padded gather to the destination rank
  HostSharedBuffer   single-node delivery: every rank copies its rows device -> host shared memory over its own
                     PCIe link, the destination rank maps the same segment (no rank-0 D2H of everybody's samples)

Why a shard is not simply "its own batch": the reference runs ONE padded batch (models/fastpitch/networks.py:140-195) and
FastPitch is not batch-invariant — PositionwiseConvFF and the predictors stack convolutions without a mask in between
(transformer.py:83-85, model.py:129-133), so an utterance's last valid positions depend on whether at least one padded
position follows them (any number >= 1 gives the same values: padded inputs are exactly zero). In a shard the longest
utterance would lose that padded position, so `plan_shards` gives every shard whose longest utterance is shorter than the
global maximum ONE extra padded token column, and `synthesize` does the same in the frame domain with one 4-byte
all-reduce(max) of the frame count between the duration stage and the decoder — the only data-dependent exchange of the
path. With both, the sharded result equals the single-batch result bit for bit (tests/test_gpu_parallel.py).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Indices per rank. Sort by length (descending, stable), deal in a snake (0..N-1, N-1..0, ...)
    so that the summed length — a proxy for frames, hence cost — is balanced."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order):
        rnd, off = divmod(pos, world_size)
        rank = off if rnd % 2 == 0 else world_size - 1 - off
        shards[rank].append(idx)
    return shards


def plan_shards(lengths: Sequence[int], world_size: int) -> Tuple[List[List[int]], List[int]]:
    """(shards, pad_to): shards as `shard_utterances`; pad_to[r] = token columns rank r pads its batch to, i.e. its own
    maximum length plus one column when that maximum is below the global one (see the module docstring)."""
    shards = shard_utterances(lengths, world_size)
    l_max = max(int(n) for n in lengths) if len(lengths) else 0
    pad_to = []
    for idxs in shards:
        own = max((int(lengths[i]) for i in idxs), default=0)
        pad_to.append(own + 1 if 0 < own < l_max else own)
    return shards, pad_to


class HostSharedBuffer:
    """A POSIX shared-memory segment viewed as an fp32 [rows, cols] tensor by every rank of ONE node, page-locked in each
    process so that device -> host copies into it run at full PCIe speed from every GPU at once."""

    def __init__(self, name: str, rows: int, cols: int, create: bool):
        from multiprocessing import shared_memory
        nbytes = max(4, rows * cols * 4)
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes if create else 0)
        if not create:
            # Python < 3.13 registers an ATTACHED segment with this process's resource tracker too, which then tries to
            # unlink the creator's segment at exit ("leaked shared_memory objects" warnings): only the creator owns it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        import numpy as np
        self.rows, self.cols = rows, cols
        self.array = np.ndarray((rows, cols), dtype=np.float32, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self.registered = False
        if torch.cuda.is_available():
            r = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.tensor.numel() * 4, 0)
            self.registered = int(r) == 0
        self.owner = create

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False
        self.tensor = None
        self.array = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


# Persistent delivery segments. Creating a shared segment and page-locking it (cudaHostRegister) in every rank costs
# hundreds of milliseconds per GiB — several times the synthesis itself when done per call (measured: 438 ms per step at
# N = 2 against 122 ms of compute). Two segments per (group, destination) are kept and used alternately, grown on demand:
# the waveforms a call returns are VIEWS of one of them and stay valid until the call after next (clone to keep longer).
_SHM_POOL = {}


def _shm_segment(rank: int, dst: int, group, n_floats: int, dev) -> HostSharedBuffer:
    """Collective over `group`: every rank passes the same size and makes the same reuse / regrow decision. The segment is
    used as a flat fp32 array (`.tensor[0]`); rank r's rows are a CONTIGUOUS [B_r, cols_r] block of it, so the device ->
    host copies are plain contiguous transfers."""
    key = (id(group) if group is not None else 0, dst)
    pool = _SHM_POOL.setdefault(key, {'bufs': [None, None], 'turn': 0})
    turn = pool['turn']
    pool['turn'] ^= 1
    buf = pool['bufs'][turn]
    if buf is not None and buf.cols >= n_floats:
        return buf
    if buf is not None:
        dist.barrier(group=group)            # nobody still copies into the old segment
        buf.close()
    cap = max(1, -(-n_floats // (1 << 20)) * (1 << 20))
    name_t = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == dst:
        import os
        name_t[0] = int.from_bytes(os.urandom(6), 'little')
    dist.broadcast(name_t, src=dst, group=group)
    name = 'ttsb_%x' % int(name_t[0])
    ok = torch.ones(1, dtype=torch.int64, device=dev)
    buf = None
    if rank == dst:
        try:
            import os
            if os.environ.get('TTSB_SHM_DISABLE') == '1':          # tests: exercise the fallback
                raise OSError('shared segment disabled')
            buf = HostSharedBuffer(name, 1, cap, create=True)
        except Exception:                    # e.g. /dev/shm too small: every rank falls back to the NCCL delivery
            ok[0] = 0
    dist.broadcast(ok, src=dst, group=group)  # (also orders creation before the attach below)
    if int(ok[0]) == 0:
        pool['bufs'][turn] = None
        return None
    if rank != dst:
        buf = HostSharedBuffer(name, 1, cap, create=False)
    pool['bufs'][turn] = buf
    return buf


def _close_shm_pool():
    for pool in _SHM_POOL.values():
        for b in pool['bufs']:
            if b is not None:
                b.close()
    _SHM_POOL.clear()


import atexit  # noqa: E402
atexit.register(_close_shm_pool)


def _synthesize_host_shm(model, id_list, shards, pad_to, rank, world, dst, group, frame_len_hook, speed, speaker_id, denoise,
                         pitch_transform, max_duration, return_stats):
    """deliver = 'host_shm': every rank's generator writes its finished utterance groups straight into its rows of the
    shared pinned segment (device -> host copies under the next group's compute, each over the rank's own PCIe link);
    `dst` then only needs every rank's sample counts and sort permutation (two small all_gathers, which also order the
    copies before the read) to hand out views in input order."""
    dev = model.device
    mine = shards[rank]
    state = {}

    def host_alloc(b_local: int, n_cols: int):
        meta = torch.tensor([b_local, n_cols], dtype=torch.int64, device=dev)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        rows_of = [int(m[0]) for m in metas]
        cols_of = [int(m[1]) for m in metas]
        sizes = [r * c for r, c in zip(rows_of, cols_of)]
        buf = _shm_segment(rank, dst, group, sum(sizes), dev)
        off = sum(sizes[:rank])
        state.update(buf=buf, rows_of=rows_of, cols_of=cols_of, sizes=sizes)
        if buf is None:
            return None                      # no shared segment: the waveforms stay on the device, NCCL delivers them
        return buf.tensor[0, off:off + b_local * n_cols].view(b_local, n_cols) if b_local else None

    if mine:
        wav, n_samples, inverse, _ = model.synthesize_ids([id_list[i] for i in mine], speed, speaker_id, denoise,
                                                          pitch_transform, max_duration, to_cpu=False,
                                                          pad_to=pad_to[rank], frame_len_hook=frame_len_hook,
                                                          return_padded=True, host_alloc=host_alloc)
        info = torch.stack([n_samples.to(dev), inverse.to(dev)], dim=1).contiguous()     # [B_local, 2], sorted-row order
    else:
        frame_len_hook(0)                     # keep the collectives matched
        host_alloc(0, 1)
        wav = torch.zeros(0, 1, dtype=torch.float32, device=dev)
        n_samples = torch.zeros(0, dtype=torch.int64, device=dev)
        inverse = torch.zeros(0, dtype=torch.int64)
        info = torch.zeros(0, 2, dtype=torch.int64, device=dev)
    stats = {'frames': int(n_samples.sum()) // max(1, model.vocoder.hop), 'utterances': len(mine)}
    if state['buf'] is None:
        # fallback (the shared segment could not be created): padded NCCL gather + one device -> host copy on dst
        rows = inverse.to(dev)
        got = gather_waveforms(wav.index_select(0, rows), n_samples.index_select(0, rows), dst=dst, group=group)
        res = None
        if rank == dst:
            per_rank = []
            for w, c in zip(*got):
                host = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
                host.copy_(w, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                per_rank.append([host[k, :int(n)] for k, n in enumerate(c.tolist())])
            res = unshard(per_rank, shards)
        return (res, stats) if return_stats else res
    rows_of = state['rows_of']
    infos = [torch.empty(r * 2, dtype=torch.int64, device=dev) for r in rows_of]
    if len(set(rows_of)) == 1:
        dist.all_gather(infos, info.reshape(-1), group=group)
    else:
        _all_gather_ragged(infos, info.reshape(-1), group)
    res = None
    if rank == dst:
        buf = state['buf']
        per_rank, at = [], 0
        for r in range(world):
            inf = infos[r].reshape(-1, 2).tolist()
            block = buf.tensor[0, at:at + state['sizes'][r]].view(rows_of[r], max(1, state['cols_of'][r])) if rows_of[r] else None
            # shard position j of rank r sits in sorted row inverse[j]
            per_rank.append([block[int(inf[j][1]), :int(inf[int(inf[j][1])][0])] for j in range(rows_of[r])])
            at += state['sizes'][r]
        res = unshard(per_rank, shards)
    return (res, stats) if return_stats else res


@torch.inference_mode()
def synthesize(model, id_list: List[torch.Tensor], speed=1., speaker_id=0, denoise=0., pitch_transform=None,
               max_duration=75, dst: int = 0, group=None, deliver: str = 'nccl', return_stats: bool = False):
    """FastPitch2Wave.synthesize_ids over the ranks of `group`: every rank passes the SAME utterance list (token ids are
    tiny), synthesizes its own shard and delivers the waveforms to rank `dst`, which returns the list of 1-D waveforms in
    input order (None elsewhere). deliver = 'nccl' (padded gather over NCCL/NVLink; rows stay on dst's device),
    'nccl_host' (the same + one device -> host copy on dst), 'host_shm' (single node: every rank copies its rows into one
    shared pinned host segment over its own PCIe link; dst returns CPU tensors)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(x.numel()) for x in id_list]
    shards, pad_to = plan_shards(lengths, world)
    mine = shards[rank]
    dev = model.device

    def frame_len_hook(t_local: int) -> int:
        # the frame-domain twin of pad_to: one extra padded frame when this shard's longest mel is shorter than the
        # global longest (4-byte all-reduce; the reference's single batch knows the global maximum by construction)
        if world == 1:
            return t_local
        t = torch.tensor([t_local], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return t_local + 1 if t_local < int(t[0]) else t_local

    if world == 1 and deliver != 'nccl':
        # one rank, host delivery: FastPitch2Wave.synthesize_ids itself (its device -> host copies run under the generator)
        wavs, _ = model.synthesize_ids(id_list, speed, speaker_id, denoise, pitch_transform, max_duration, to_cpu=True)
        stats = {'frames': sum(int(w.numel()) for w in wavs) // max(1, model.vocoder.hop), 'utterances': len(wavs)}
        return (wavs, stats) if return_stats else wavs
    if world > 1 and deliver == 'host_shm':
        return _synthesize_host_shm(model, id_list, shards, pad_to, rank, world, dst, group, frame_len_hook, speed, speaker_id,
                                    denoise, pitch_transform, max_duration, return_stats)
    if mine:
        wav, n_samples, inverse, _ = model.synthesize_ids([id_list[i] for i in mine], speed, speaker_id, denoise,
                                                          pitch_transform, max_duration, to_cpu=False,
                                                          pad_to=pad_to[rank], frame_len_hook=frame_len_hook,
                                                          return_padded=True)
        # sorted-batch row of the shard's k-th utterance
        rows = inverse.to(dev)
        wav = wav.index_select(0, rows)
        n_samples = n_samples.index_select(0, rows)
    else:
        frame_len_hook(0)                     # keep the collective matched
        wav = torch.zeros(0, 1, dtype=torch.float32, device=dev)
        n_samples = torch.zeros(0, dtype=torch.int64, device=dev)
    stats = {'frames': int(n_samples.sum()) // max(1, model.vocoder.hop), 'utterances': len(mine)}

    if world == 1:
        if deliver != 'nccl':
            # one D2H copy of the padded batch into pinned memory (the same staging FastPitch2Wave.synthesize_ids uses)
            host = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
            host.copy_(wav, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            wav = host
        out = [wav[k, :int(n)] for k, n in enumerate(n_samples.tolist())]
        res = unshard([out], shards)
        return (res, stats) if return_stats else res

    got = gather_waveforms(wav, n_samples, dst=dst, group=group)
    res = None
    if rank == dst:
        ws, cs = got
        per_rank = []
        for w, c in zip(ws, cs):
            if deliver == 'nccl_host':
                host = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
                host.copy_(w, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                w = host
            per_rank.append([w[k, :int(n)] for k, n in enumerate(c.tolist())])
        res = unshard(per_rank, shards)
    return (res, stats) if return_stats else res


def _all_gather_ragged(outs: List[torch.Tensor], mine: torch.Tensor, group=None):
    """all_gather for per-rank tensors of different lengths (pad to the longest, trim on arrival)."""
    n_max = max(int(o.numel()) for o in outs)
    pad = torch.zeros(n_max, dtype=mine.dtype, device=mine.device)
    pad[:mine.numel()] = mine
    bufs = [torch.empty_like(pad) for _ in outs]
    dist.all_gather(bufs, pad, group=group)
    for o, b in zip(outs, bufs):
        o.copy_(b[:o.numel()])


def gather_waveforms(wav: torch.Tensor, n_samples: torch.Tensor, dst: int = 0,
                     group=None) -> Optional[Tuple[List[torch.Tensor], List[torch.Tensor]]]:
    """wav [B_local, N_local] padded waveforms, n_samples [B_local] valid samples per row.
    Returns on `dst`: (list over ranks of [B_r, N_max] tensors, list over ranks of [B_r] counts);
    None elsewhere. Works for ragged B_local / N_local across ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = wav.device
    meta = torch.tensor([wav.shape[0], wav.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    b_max = int(max(int(m[0]) for m in metas))
    n_max = int(max(int(m[1]) for m in metas))
    pad = torch.zeros(b_max, n_max, dtype=wav.dtype, device=dev)
    pad[:wav.shape[0], :wav.shape[1]] = wav
    cnt = torch.zeros(b_max, dtype=torch.int64, device=dev)
    cnt[:wav.shape[0]] = n_samples.to(dev, torch.int64)
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        cnts = [torch.empty_like(cnt) for _ in range(world)]
    else:
        bufs = cnts = None
    dist.gather(pad, bufs, dst=dst, group=group)
    dist.gather(cnt, cnts, dst=dst, group=group)
    if rank != dst:
        return None
    out_w, out_c = [], []
    for r in range(world):
        b = int(metas[r][0])
        out_w.append(bufs[r][:b])
        out_c.append(cnts[r][:b])
    return out_w, out_c


def unshard(per_rank_items: List[list], shards: List[List[int]]) -> list:
    """Inverse of shard_utterances for per-rank result lists."""
    n = sum(len(s) for s in shards)
    out = [None] * n
    for items, idxs in zip(per_rank_items, shards):
        for item, i in zip(items, idxs):
            out[i] = item
    return out
