"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding rule and the
collectives the product calls (bayes_kit_b200.dist): the Stretcher's complementary-half
all-gather with known shard sizes, the shared-seed broadcast, and the two-pass all-reduce of
R-hat moment sums.  (The sharded SMC uses no host-side collective: its cross-rank protocol runs
inside the kernels and is covered by tests/test_gpu_sharded_smc.py and scripts/r2/dist_check.py.)
The oracle plays the single-process reference the sharded result must match."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import diagnostics as od


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bayes_kit_b200 import dist
    try:
        assert dist.rank_world() == (rank, world)
        # ragged all-gather (SMC particles / log-weights; R-hat moments)
        M = 11
        lo, hi = dist.shard_range(M, rank, world)
        full = torch.arange(M * 3, dtype=torch.float64).reshape(M, 3)
        got = dist.all_gather_cat(full[lo:hi].clone())
        assert torch.equal(got, full)
        # the same gather when every rank knows the shard sizes (no size exchange, no host sync)
        got = dist.all_gather_cat(full[lo:hi].clone(), sizes=dist.shard_sizes(M, world))
        assert torch.equal(got, full)
        half = torch.arange(8.0).reshape(4, 2)       # equal shards: all_gather_into_tensor path
        l2, h2 = dist.shard_range(4, rank, world)
        assert torch.equal(dist.all_gather_cat(half[l2:h2].clone(), sizes=dist.shard_sizes(4, world)), half)
        # seed=None: every rank must end up with rank 0's Philox key
        seed = dist.shared_seed(1000 + rank)
        assert seed == 1000
        # cross-chain R-hat as the two-pass all-reduce of per-parameter sums the product performs
        # (rhat._rhat_allreduce: {n, sum mean, sum var} -> grand mean -> sum sq dev; 4 doubles / parameter)
        rng = np.random.default_rng(0)
        ch = rng.normal(size=(7, 50, 3)) + 1e6 + rng.normal(size=(7, 1, 3))     # large offset: cancellation trap
        lo, hi = dist.shard_range(7, rank, world)
        m = torch.as_tensor(ch[lo:hi].mean(1))
        v = torch.as_tensor(ch[lo:hi].var(1, ddof=1))
        sums = torch.stack([torch.full((3,), float(hi - lo), dtype=torch.float64), m.sum(0), v.sum(0)], 1)
        tdist.all_reduce(sums)
        ref = sums[:, 1] / sums[:, 0]
        sq = ((m - ref) ** 2).sum(0)
        tdist.all_reduce(sq)
        r = torch.sqrt((50.0 - 1) / 50.0 + (sq / (sums[:, 0] - 1)) / (sums[:, 2] / sums[:, 0]))
        for p_ in range(3):
            assert abs(float(r[p_]) - od.rhat(list(ch[:, :, p_]))) < 1e-9
        out_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out_q.put((rank, repr(e)))
    finally:
        tdist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=100) for _ in ps]
    for p in ps:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
