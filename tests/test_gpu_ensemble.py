"""Stretcher (ensemble.py:9-66; upstream implementation commented out -> PARITY UNPINNED by the
reference): the device stretch move against the oracle restatement under injected uniforms, the
constructor's validation messages from the reference's sketch, and a posterior check."""
import numpy as np
import pytest
import torch

from _dev import np_
from oracle import samplers as osm
from oracle.models import DensePrecGauss, DiagGauss, IsoGauss

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,D,W", [("iso", 5, 16), ("diag", 37, 80), ("dense", 64, 130), ("iso", 1, 4)])
def test_stretch_matches_oracle(bk, kind, D, W):
    rng = np.random.default_rng(D + W)
    if kind == "iso":
        om, dm = IsoGauss(D, 1.5), bk.IsoGauss(D, 1.5, dtype=torch.float64)
    elif kind == "diag":
        mu, pr = rng.normal(size=D), rng.uniform(0.5, 3, D)
        om, dm = DiagGauss(mu, pr), bk.DiagGauss(mu, pr, dtype=torch.float64)
    else:
        P = DensePrecGauss.c2_precision(D, 1)
        om, dm = DensePrecGauss(P), bk.DensePrecGauss(P, dtype=torch.float64)
    th0 = rng.normal(size=(W, D))
    n = 6
    us = rng.random((n, W, 3))
    want, wacc = osm.stretch(om, th0, us, a=2.0)
    s = bk.Stretcher(dm, a=2.0, walkers=W, init=th0)
    for t in range(n):
        got = s.sample(uniforms=us[t])
        assert np.array_equal(np_(s.last_accept).astype(bool), wacc[t])
        np.testing.assert_allclose(np_(got), want[t], rtol=1e-10, atol=1e-10)
    assert 0.2 < wacc.mean() < 1.0


def test_stretcher_validation(bk):
    m = bk.IsoGauss(3)
    with pytest.raises(ValueError, match="stretch bound must be greater than or equal to 1"):
        bk.Stretcher(m, a=0.5)
    for w in (0, -2, 7):
        with pytest.raises(ValueError, match="walkers must be strictly positive, even integer"):
            bk.Stretcher(m, walkers=w)
    with pytest.raises(ValueError, match="init must be shape of draw"):
        bk.Stretcher(m, walkers=8, init=np.zeros((8, 2)))
    s = bk.Stretcher(m)                        # defaults: a = 2, walkers = 2 * dims (ensemble.py:33)
    assert s.sample().shape == (6, 3)
    assert next(iter(s)).shape == (6, 3)


def test_stretcher_posterior(bk):
    """fp32 device-Philox: the ensemble reproduces a correlated Gaussian's moments."""
    D = 8
    P = DensePrecGauss.c2_precision(D, 2)
    cov = np.linalg.inv(P)
    s = bk.Stretcher(bk.DensePrecGauss(P), walkers=4096, seed=5)
    for _ in range(300):
        s.sample()
    acc = float(s.last_accept.float().mean())
    assert 0.2 < acc < 0.9
    d = np_(s.sample_n(40)).reshape(-1, D).astype(np.float64)
    n_eff = d.shape[0] / 40.0                    # sweeps are strongly autocorrelated
    assert np.all(np.abs(d.mean(0)) <= 5 * np.sqrt(np.diag(cov) / n_eff) + 1e-3)
    assert np.all(np.abs(d.var(0) / np.diag(cov) - 1) <= 5 * np.sqrt(2 / n_eff) + 1e-2)
