"""Peer-accessible device memory for the multi-GPU paths (plumbing only).

One process per GPU.  The sharded SMC (and the Stretcher's complementary half)
exchange data inside their kernels through memory every rank can address over
NVLink: torch symmetric memory provides the allocation and the table of peer
pointers (``torch.distributed._symmetric_memory``; one rendezvous per buffer
set, at construction time only).  ``FakeWorld`` gives the same tables for G
"ranks" that all live on ONE device -- the single-GPU tests loop the ranks of a
sharded run in lockstep through the very same kernels (SURVEY 4 "fake world").
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import dist as D_


class FakeWorld:
    """G logical ranks on one device.  ``world.rank(r)`` is passed as ``group=``."""

    def __init__(self, world: int):
        self.world = int(world)
        self._regions: Dict[str, List] = {}

    def rank(self, r: int) -> "FakeRank":
        return FakeRank(self, int(r))


class FakeRank:
    def __init__(self, world: FakeWorld, rank: int):
        self.fake, self.rank, self.world = world, rank, world.world


def rank_world(group=None) -> Tuple[int, int]:
    if isinstance(group, FakeRank):
        return group.rank, group.world
    return D_.rank_world(group)


class PeerRegion:
    """A set of equally laid-out byte regions, one per rank, each addressable by every
    rank.  ``sizes``: name -> bytes; ``tensor(name, dtype, shape)`` views the LOCAL
    region, ``ptrs(name)`` lists every rank's base address of that field."""

    def __init__(self, key: str, sizes: Dict[str, int], device, group=None):
        self.rank, self.world = rank_world(group)
        self.device = device
        self._off, off = {}, 0
        for name, nbytes in sizes.items():
            self._off[name] = off
            off += (int(nbytes) + 255) // 256 * 256
        self.nbytes = max(off, 256)
        self._handle = None
        if isinstance(group, FakeRank):
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
            reg = group.fake._regions.setdefault(key, [None] * self.world)
            reg[self.rank] = self.buf
            self._fake_reg = reg
            self._bases = None
        elif self.world == 1:
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
            self._bases = [self.buf.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem
            with torch.cuda.device(device):
                self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
                self.buf.zero_()
                self._handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self._bases = [int(p) for p in self._handle.buffer_ptrs]
            torch.cuda.synchronize(device)
            dist.barrier(group=group)     # every rank's region is zeroed before anybody posts into it

    def bases(self) -> Sequence[int]:
        if self._bases is None:
            reg = self._fake_reg
            if any(b is None for b in reg):
                raise RuntimeError("FakeWorld: construct every rank before running a step")
            self._bases = [b.data_ptr() for b in reg]
        return self._bases

    def ptrs(self, name: str) -> List[int]:
        return [b + self._off[name] for b in self.bases()]

    def tensor(self, name: str, dtype: torch.dtype, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        o = self._off[name]
        return self.buf[o:o + nbytes].view(dtype).view(*shape)
