"""Multi-GPU layer: utterances are independent end to end, so a batch shards across ranks with no
data-path collective; the only exchange is the final delivery of waveforms.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).
  shard_utterances   length-balanced "snake" deal of a length-sorted utterance list
  gather_waveforms   all_gather of sample counts, then one padded gather to the destination rank
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
