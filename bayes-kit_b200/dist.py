"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests).  Sampling is communication-free (chains shard
contiguously, Philox is keyed by the GLOBAL chain id); collectives appear only
at the naturally global steps: the Stretcher's complementary half (all-gather) and
cross-chain R-hat moment sums (all-reduce); the sharded SMC talks through peer
memory inside its kernels (peer.py) and calls no collective per temperature.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; the first n % world
    ranks get one extra item."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def all_gather_cat(t: torch.Tensor, group=None, sizes=None) -> torch.Tensor:
    """Concatenate every rank's tensor along dim 0.  ``sizes``: every rank's first dimension when
    the caller knows them (shards of a known total: ``shard_sizes``) -- then nothing is exchanged
    but the data and there is no host synchronisation; without it the sizes are all-gathered first
    (one blocking read)."""
    rank, world = rank_world(group)
    if world == 1:
        return t
    if sizes is None:
        n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        sizes = [int(s.item()) for s in got]
    sizes = [int(s) for s in sizes]
    if sizes[rank] != t.shape[0]:
        raise ValueError(f"all_gather_cat: rank {rank} holds {t.shape[0]} rows, sizes says {sizes[rank]}")
    mx = max(sizes)
    if min(sizes) == mx:
        out = torch.empty((mx * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    pad = t
    if t.shape[0] < mx:
        pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)]).contiguous()


def shared_seed(seed: int, group=None, device=None) -> int:
    """``seed=None`` resolves to fresh entropy on every rank; the ranks of one sampler must share the
    Philox key, so rank 0's value is broadcast (one-off, at construction)."""
    rank, world = rank_world(group)
    if world == 1:
        return seed
    t = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64,
                     device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(t.item())
