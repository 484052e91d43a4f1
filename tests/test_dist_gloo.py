"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding and
the three collective call sites (SMC normaliser / resample all-gather, R-hat
moment all-gather) of bayes_kit_b200.dist.  The oracle plays the single-process
reference the sharded result must match."""
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
        # log-sum-exp normaliser from per-rank (max, sum exp) pairs
        rng = np.random.default_rng(0)
        logw = torch.as_tensor(rng.normal(size=M) * 30)
        loc = logw[lo:hi]
        lmax = loc.max().reshape(1)
        lsum = torch.exp(loc - lmax).sum().reshape(1)
        gmax, gsum = dist.all_reduce_logsumexp(lmax, lsum)
        want = torch.logsumexp(logw, 0)
        assert abs(float(gmax + torch.log(gsum)) - float(want)) < 1e-12
        # cross-chain R-hat from sharded per-chain moments == oracle on all chains
        ch = rng.normal(size=(6, 50)) + rng.normal(size=(6, 1))
        lo, hi = dist.shard_range(6, rank, world)
        m = torch.as_tensor(ch[lo:hi].mean(1)).reshape(-1, 1)
        v = torch.as_tensor(ch[lo:hi].var(1, ddof=1)).reshape(-1, 1)
        m, v = dist.all_gather_cat(m), dist.all_gather_cat(v)
        nbar = 50.0
        r = float(torch.sqrt((nbar - 1) / nbar + m.var(0, unbiased=True) / v.mean(0)))
        assert abs(r - od.rhat(list(ch))) < 1e-12
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
