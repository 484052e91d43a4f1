"""tcgen05 path of the dense-precision Gaussian (fp32 timed mode) against fp64
NumPy / the oracle.  Tolerances: the endpoint gradient uses a 3-pass bf16
split (~2^-16 relative per product, fp32 accumulate); interior leapfrog
gradients use plain bf16 operands (2^-9), so multi-step trajectories agree
with the fp64 oracle to ~1e-3 while the energies that enter the accept test
stay fp32-accurate."""
import numpy as np
import pytest
import torch

from _dev import np_
from oracle import samplers as osm
from oracle.models import DensePrecGauss

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,C,with_mu", [(1000, 300, False), (200, 70, True), (128, 256, False), (64, 5, True)])
def test_gradient_split_precision(bk, D, C, with_mu):
    rng = np.random.default_rng(D + C)
    P = DensePrecGauss.c2_precision(D, 2)
    mu = rng.normal(size=D) if with_mu else None
    th = rng.normal(size=(C, D))
    m = bk.DensePrecGauss(P, mu, dtype=torch.float32)
    lp, g = m.log_density_gradient(th)
    om = DensePrecGauss(P, mu)
    want = [om.log_density_gradient(t) for t in th.astype(np.float32).astype(np.float64)]
    wg = np.stack([w[1] for w in want]); wl = np.array([w[0] for w in want])
    scale = np.abs(wg).max()
    assert np.abs(np_(g) - wg).max() <= 2e-4 * scale      # bf16x3 + fp32 accumulation over D terms
    np.testing.assert_allclose(np_(lp), wl, rtol=2e-4, atol=1e-3)


def _bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def _emulated_hmc_draw(P, th, z, eps, L):
    """fp64 restatement of the tensor-core pipeline's arithmetic: accurate
    gradients at both trajectory ends, bf16-rounded operands in between."""
    Pb = _bf16(P)
    g = -(th @ P)
    r = z - 0.5 * eps * g + eps * g
    q = th + eps * r
    for _ in range(1, L):
        g = -(_bf16(q) @ Pb)
        r = r + eps * g
        q = q + eps * r
    g = -(q @ P)
    r = r + 0.5 * eps * g
    h0 = -0.5 * np.einsum("cd,cd->c", th @ P, th) - 0.5 * np.einsum("cd,cd->c", z, z)
    h1 = -0.5 * np.einsum("cd,cd->c", q @ P, q) - 0.5 * np.einsum("cd,cd->c", r, r)
    return q, h0, h1


@pytest.mark.parametrize("D,C,L,eps", [(1000, 300, 1, 0.2), (1000, 300, 4, 0.1), (200, 513, 10, 0.1),
                                       (128, 64, 3, 0.2), (1000, 257, 10, 0.1)])
def test_hmc_tc_one_draw(bk, D, C, L, eps):
    """One draw under injected streams.  (i) against the fp64 emulation of the
    pipeline's own arithmetic: proposals within 1e-3, energies within 2e-2,
    identical accept decisions away from ties; (ii) against the plain fp64
    oracle: the bf16 interior gradients move the trajectory by O(1e-2) at most
    and do not change the acceptance behaviour."""
    rng = np.random.default_rng(L * 7 + D)
    P = DensePrecGauss.c2_precision(D, 0)
    th0 = rng.normal(size=(C, D)).astype(np.float32).astype(np.float64)
    zs = rng.standard_normal((1, C, D)).astype(np.float32).astype(np.float64)
    us = rng.random((1, C)).astype(np.float32).astype(np.float64)
    s = bk.HMCDiag(bk.DensePrecGauss(P, dtype=torch.float32), eps, L, init=th0)
    d, l = s.sample_n(1, normals=zs, uniforms=us)
    d, l, a = np_(d)[0].astype(np.float64), np_(l)[0].astype(np.float64), np_(s.last_accept)[0].astype(bool)
    q, h0, h1 = _emulated_hmc_draw(P, th0, zs[0], eps, L)
    margin = np.log(us[0]) - (h1 - h0)
    want_a = margin < 0
    clear = np.abs(margin) > 0.05
    assert np.array_equal(a[clear], want_a[clear])
    # bf16 rounding-boundary flips of single operands move a few elements by O(1e-3);
    # the bulk agrees to fp32 rounding
    err = np.abs(d[a] - q[a])
    assert np.quantile(err, 0.999) <= 2e-4 and err.max() <= 2e-2, (np.quantile(err, 0.999), err.max())
    if (~a).any():
        assert np.abs(d[~a] - th0[~a]).max() == 0
    np.testing.assert_allclose(l[clear], np.where(want_a, h1, h0)[clear], rtol=0, atol=2e-2)
    od, ol, oa = osm.hmc_diag_batch(DensePrecGauss(P), th0, zs, us, eps, L)
    assert (a != oa[0]).mean() <= 0.02
    same = a == oa[0]
    assert np.abs(d[same] - od[0][same]).max() <= (2e-5 if L == 1 else 5e-3 * L)


def test_tc_matches_cuda_core_path(bk, monkeypatch):
    """L=1 has no bf16 interior step: tensor-core and CUDA-core fp32 pipelines agree closely."""
    rng = np.random.default_rng(5)
    D, C = 256, 200
    P = DensePrecGauss.c2_precision(D, 1)
    th0, zs, us = rng.normal(size=(C, D)), rng.standard_normal((3, C, D)), rng.random((3, C))
    res = []
    for off in ("0", "1"):
        monkeypatch.setenv("BK_DISABLE_TC", off)
        s = bk.HMCDiag(bk.DensePrecGauss(P, dtype=torch.float32), 0.2, 1, init=th0)
        d, l = s.sample_n(3, normals=zs, uniforms=us)
        res.append((np_(d), np_(l), np_(s.last_accept)))
    same = (res[0][2] == res[1][2]).all(0)
    assert same.mean() > 0.97
    assert np.abs(res[0][0][:, same] - res[1][0][:, same]).max() < 1e-4
    assert np.abs(res[0][1][:, same] - res[1][1][:, same]).max() < 5e-3


def test_tc_posterior_moments(bk):
    D = 256
    P = DensePrecGauss.c2_precision(D, 3)
    var = np.diag(np.linalg.inv(P))
    # eps * L = 1.0 keeps omega*T < pi for every eigen-frequency (sqrt(lambda) <= 2.24):
    # fixed-length HMC is non-ergodic for modes with omega*T = pi (x -> -x exactly)
    s = bk.HMCDiag(bk.DensePrecGauss(P), 0.1, 10, chains=2048, seed=3)
    s.sample_n(60, keep_draws=False)
    draws, _ = s.sample_n(200)
    acc = float(s.last_accept.float().mean())
    assert 0.8 < acc <= 1.0
    d = np_(draws).reshape(-1, D).astype(np.float64)
    n_eff = 2048 * 200 / 4.0   # conservative: |lag-1 autocorrelation| of x and x^2 is < 0.6
    assert np.all(np.abs(d.mean(0)) <= 4 * np.sqrt(var / n_eff) + 1e-3)
    assert np.all(np.abs(d.var(0, ddof=1) - var) <= 4 * var * np.sqrt(2 / n_eff) + 1e-3)


def test_logreg_tc_gradient_and_hmc(bk):
    """tcgen05 path of the hierarchical logistic regression plugin (interior leapfrog
    gradients, bf16 operands): gradient within bf16 tolerance of the fp64 numpy model;
    an fp32 HMC draw (tensor-core interior steps + fp32 CUDA-core endpoint) lands within
    the trajectory tolerance of the fp64 oracle with the same accept decisions."""
    from oracle.models import HierLogReg
    rng = np.random.default_rng(0)
    for (N, Dx, C) in [(1000, 37, 70), (5000, 100, 300), (129, 5, 3)]:
        X, y = HierLogReg.c3_data(N, Dx, seed=1)
        om = HierLogReg(X, y)
        th = (rng.normal(size=(C, Dx + 2)) * 0.4).astype(np.float32).astype(np.float64)
        dm = bk.HierLogReg(X, y, dtype=torch.float32)
        lp, g = dm.log_density_gradient(th, fast=True)
        want = [om.log_density_gradient(t) for t in th]
        wl, wg = np.array([w[0] for w in want]), np.stack([w[1] for w in want])
        assert np.abs(np_(g) - wg).max() <= 1e-2 * np.abs(wg).max(), np.abs(np_(g) - wg).max() / np.abs(wg).max()
        np.testing.assert_allclose(np_(lp), wl, rtol=5e-3, atol=0.5)
    N, Dx, C, L, eps = 5000, 100, 200, 8, 0.01
    X, y = HierLogReg.c3_data(N, Dx, seed=2)
    th0 = (rng.normal(size=(C, Dx + 2)) * 0.2).astype(np.float32).astype(np.float64)
    zs = rng.standard_normal((1, C, Dx + 2)).astype(np.float32).astype(np.float64)
    us = rng.random((1, C)).astype(np.float32).astype(np.float64)
    od, ol, oa = osm.hmc_diag_batch(HierLogReg(X, y), th0, zs, us, eps, L)
    s = bk.HMCDiag(bk.HierLogReg(X, y, dtype=torch.float32), eps, L, init=th0)
    d, l = s.sample_n(1, normals=zs, uniforms=us)
    a = np_(s.last_accept)[0].astype(bool)
    assert (a != oa[0]).mean() <= 0.03
    same = a == oa[0]
    assert np.abs(np_(d)[0][same] - od[0][same]).max() <= 2e-2


@pytest.mark.parametrize("D,C,eps", [(1000, 300, 2e-3), (200, 513, 1e-2), (128, 64, 5e-3)])
def test_mala_tc_vs_oracle(bk, D, C, eps):
    """MALA on the dense plugin (fp32) runs as begin -> split-precision tcgen05 gradient ->
    accept (one leapfrog step of size sqrt(2 eps), test_equivalencies.py:12-32).  Against the
    fp64 oracle's MALA (mala.py:40-66) under the same injected streams: identical accept
    decisions away from ties, draws to fp32 rounding, log p(theta) to the split tolerance."""
    rng = np.random.default_rng(D + 1)
    P = DensePrecGauss.c2_precision(D, 0)
    n = 3
    th0 = rng.normal(size=(C, D)).astype(np.float32).astype(np.float64)
    zs = rng.standard_normal((n, C, D)).astype(np.float32).astype(np.float64)
    us = rng.random((n, C)).astype(np.float32).astype(np.float64)
    od, ol, oa = osm.mala_batch(DensePrecGauss(P), th0, zs, us, eps)
    s = bk.MALA(bk.DensePrecGauss(P, dtype=torch.float32), eps, init=th0)
    d, l = s.sample_n(n, normals=zs, uniforms=us)
    d, l, a = np_(d).astype(np.float64), np_(l).astype(np.float64), np_(s.last_accept).astype(bool)
    same = np.ones(C, dtype=bool)
    for t in range(n):
        same &= a[t] == oa[t]                    # a chain is compared until its first flipped decision
        assert same.mean() >= 0.97
        assert np.abs(d[t][same] - od[t][same]).max() <= 2e-5 * (t + 1)
        np.testing.assert_allclose(l[t][same], ol[t][same], rtol=2e-5, atol=2e-2)
