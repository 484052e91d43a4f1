"""Cross-rank R-hat and the sharded Stretcher on ONE GPU ("fake world"): the ranks of a multi-GPU run
are looped in one process and the collective is replaced by a host-side reduction / concatenation of the ranks'
device tensors.  Sharded results must equal the oracle (rhat.py:163-170; ensemble.py:55-63)."""
import threading

import numpy as np
import pytest
import torch

from _dev import np_
from oracle import diagnostics as od
from oracle import samplers as osm
from oracle.models import DiagGauss

pytestmark = pytest.mark.gpu


class FakeAllReduce:
    """SUM all-reduce between G threads of one process (fixed rank order)."""

    def __init__(self, world):
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def fn(self, rank):
        def reduce(t):
            self.slots[rank] = t.clone()
            self.bar.wait()
            total = torch.stack(self.slots).sum(0)
            self.bar.wait()
            t.copy_(total)
        return reduce


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("offset", [0.0, 1e7])
def test_sharded_rhat_equals_oracle(bk, world, offset):
    """[chains, draws, params] sharded over `world` ranks by chain: the two-pass moment all-reduce reproduces
    rhat() of all chains to 1e-9 -- also when the chain means sit on a large common offset (the cancellation
    case a one-pass sum of squares would lose)."""
    rng = np.random.default_rng(world)
    C, N, P = 37, 400, 5
    x = rng.normal(size=(C, N, P)) + rng.normal(size=(C, 1, P)) * 0.3 + offset
    want = np.array([od.rhat(list(x[:, :, p])) for p in range(P)])
    xd = torch.as_tensor(x, device="cuda")
    red = FakeAllReduce(world)
    out = [None] * world
    errs = []

    def work(r):
        try:
            lo, hi = bk.dist.shard_range(C, r, world)
            with torch.cuda.device(0):
                out[r] = bk.rhat(xd[lo:hi], reduce_fn=red.fn(r))
        except Exception as e:      # pragma: no cover
            errs.append(e)
            red.bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    single = np_(bk.rhat(xd))
    for r in range(world):
        np.testing.assert_allclose(np_(out[r]), want, rtol=1e-9, atol=0)
        np.testing.assert_allclose(np_(out[r]), single, rtol=1e-9, atol=0)
        assert np.array_equal(np_(out[r]), np_(out[0]))       # every rank holds the same answer


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stretcher_equals_oracle(bk, world):
    """Walkers of both halves sharded over `world` ranks; the complementary half is gathered before each
    half-step.  Under injected uniforms the sharded ensemble is the single-process oracle run."""
    rng = np.random.default_rng(9)
    D, W, n = 7, 26, 5
    mu, pr = rng.normal(size=D), rng.uniform(0.5, 3, D)
    om, dm = DiagGauss(mu, pr), bk.DiagGauss(mu, pr, dtype=torch.float64)
    th0 = rng.normal(size=(W, D))
    us = rng.random((n, W, 3))
    want, wacc = osm.stretch(om, th0, us, a=2.0)
    fw = bk.peer.FakeWorld(world)
    ranks = [bk.Stretcher(dm, a=2.0, walkers=W, init=th0, group=fw.rank(r)) for r in range(world)]
    h = W // 2
    for t in range(n):
        for half in (0, 1):
            for s in ranks:
                u_loc = np.concatenate([us[t, s._lo:s._hi], us[t, h + s._lo:h + s._hi]])
                s._half_step(half, u_loc)
        torch.cuda.synchronize()
        first = np.concatenate([np_(s._halves[0]) for s in ranks])
        second = np.concatenate([np_(s._halves[1]) for s in ranks])
        np.testing.assert_allclose(np.concatenate([first, second]), want[t], rtol=1e-10, atol=1e-10)
        acc = np.concatenate([np_(s.last_accept).reshape(2, -1)[0] for s in ranks] +
                             [np_(s.last_accept).reshape(2, -1)[1] for s in ranks]).astype(bool)
        assert np.array_equal(acc, wacc[t])
