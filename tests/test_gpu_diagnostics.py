"""GPU parity for autocorr / iat / ess / rhat: golden fixtures recorded from the
reference (1e-9 contract of BASELINE north_star; observed ~1e-13), the
reference's known-answer vectors, its ValueError contracts, and batched layouts."""
import numpy as np
import pytest
import torch

from _dev import np_
from conftest import golden
from oracle import diagnostics as od

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("tag", ["n37", "n1000", "n10000"])
def test_golden(bk, tag):
    z = golden("diagnostics")
    x = z["x_" + tag]
    np.testing.assert_allclose(np_(bk.autocorr(x)), z["autocorr_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.iat_ipse(x)), z["iat_ipse_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.iat_imse(x)), z["iat_imse_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.iat(x)), z["iat_imse_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.ess(x)), z["ess_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.ess_imse(x)), z["ess_" + tag], **TOL)
    np.testing.assert_allclose(np_(bk.ess_ipse(x)), z["ess_ipse_" + tag], **TOL)
    # reference surface: one 1-D host chain -> float
    e = bk.ess(x[0])
    assert isinstance(e, float) and abs(e - z["ess_" + tag][0]) <= 1e-9 * abs(e)


def test_rhat_golden(bk):
    z = golden("diagnostics")
    ch = z["rhat_chains"]                       # [chains, draws, params]
    np.testing.assert_allclose(np_(bk.rhat(ch)), z["rhat"], **TOL)
    r0 = bk.rhat([list(c) for c in ch[:, :, 0]])  # reference surface: list of chains
    assert isinstance(r0, float) and abs(r0 - z["rhat"][0]) < 1e-9
    # draws-first layout (what the samplers write) consumed in place
    np.testing.assert_allclose(np_(bk.rhat(np.ascontiguousarray(ch.transpose(1, 0, 2)), draws_first=True)),
                               z["rhat"], **TOL)
    # fp32 device input
    t32 = torch.as_tensor(ch, dtype=torch.float32, device="cuda")
    np.testing.assert_allclose(np_(bk.rhat(t32)), od.rhat_batch(np_(t32).astype(np.float64)), **TOL)


def test_layouts(bk):
    rng = np.random.default_rng(1)
    x = np.stack([[od.sample_ar1(0.5, 400, rng) for _ in range(3)] for _ in range(5)])  # [5, 3, 400]
    cdp = np.ascontiguousarray(x.transpose(0, 2, 1))   # [chains, draws, params]
    dcp = np.ascontiguousarray(x.transpose(2, 0, 1))   # [draws, chains, params]
    want = np.array([[od.ess(x[c, p]) for p in range(3)] for c in range(5)])
    np.testing.assert_allclose(np_(bk.ess(cdp)), want, **TOL)
    np.testing.assert_allclose(np_(bk.ess(dcp, draws_first=True)), want, **TOL)
    dev = torch.as_tensor(dcp, device="cuda")
    np.testing.assert_allclose(np_(bk.ess(dev, draws_first=True)), want, **TOL)
    ac = np_(bk.autocorr(cdp))
    assert ac.shape == (5, 3, 400)
    np.testing.assert_allclose(ac[2, 1], od.autocorr(x[2, 1]), **TOL)


def test_known_answers(bk):
    # test_autocorr.py:10-14
    np.testing.assert_allclose(np_(bk.autocorr(np.asarray([1, 0, 0, 0]))),
                               [1.000, -0.083, -0.167, -0.250], atol=0.001, rtol=0.001)
    # test_rhat.py:34-52 (vs the BDA3 brute force)
    c = [[1.01, 1.05, 0.98, 0.90, 1.23], [0.99, 1.00, 1.01, 1.15, 0.83],
         [0.84, 0.90, 0.94, 1.10, 0.92], [0.32, 1.81, 0.90, 0.10, 2.85]]
    for k in (2, 3, 4):
        assert abs(bk.rhat(c[:k]) - od.rhat(c[:k])) < 1e-12
    # ragged (test_rhat.py:45-52)
    rag = [[1.01, 1.05, 0.98, 0.90, 1.23], [0.99, 1.00, 1.01, 1.15, 0.83, 0.95]]
    assert abs(bk.rhat(rag) - od.rhat(rag)) < 1e-12
    # anti-correlated / tiny chains
    for ch in ([1.0, -0.5, 0.25, -0.3, 0.7], [0.1, 0.2, 0.1, 0.2], [3.0, 1.0, 2.0, 5.0, 4.0, 0.0]):
        assert abs(bk.iat_ipse(ch) - od.iat_ipse(np.asarray(ch))) < 1e-12
        assert abs(bk.iat_imse(ch) - od.iat_imse(np.asarray(ch))) < 1e-12


def test_ar1_closed_form(bk):
    # test_iat.py:29-54 / test_ess.py:14-39: IAT = (1+phi)/(1-phi) within 10%
    rng = np.random.default_rng(2)
    for phi in (-0.5, -0.3, -0.1, 0.1, 0.3):
        v = od.sample_ar1(phi, 20_000, rng)
        for f in (bk.iat, bk.iat_imse, bk.iat_ipse):
            np.testing.assert_allclose((1 + phi) / (1 - phi), f(v), rtol=0.1)
        np.testing.assert_allclose(20_000 * (1 - phi) / (1 + phi), bk.ess(v), rtol=0.1)


def test_exceptions(bk):
    # test_autocorr.py:42-47, test_iat.py:57-65, test_ess.py:42-50, test_rhat.py:55-68
    for bad in ([], [1.1]):
        with pytest.raises(ValueError):
            bk.autocorr(bad)
    bk.autocorr([1.1, 1.2])
    for n in range(4):
        v = np.arange(n, dtype=float)
        for f in (bk.iat, bk.iat_imse, bk.iat_ipse, bk.ess, bk.ess_imse, bk.ess_ipse):
            with pytest.raises(ValueError):
                f(v)
    for bad in ([], [[1.01, 1.2, 1.3, 1.4]], [[1, 2, 3], [4], [5, 6, 7, 8, 9]]):
        with pytest.raises(ValueError):
            bk.rhat(bad)


@pytest.mark.parametrize("mode", ["direct", "fft"])
def test_autocorr_both_paths(bk, mode, monkeypatch):
    """The direct-sum and the hand-written FFT kernels are the same estimator."""
    monkeypatch.setenv("BK_ACF", mode)
    z = golden("diagnostics")
    for tag in ("n37", "n1000") + (("n10000",) if mode == "fft" else ()):
        np.testing.assert_allclose(np_(bk.autocorr(z["x_" + tag])), z["autocorr_" + tag], **TOL)
    rng = np.random.default_rng(4)
    for n in (2, 3, 5, 64, 257, 4097):   # FFT sizes 4 .. 16384 (both sides of the in-smem limit)
        x = np.stack([od.sample_ar1(0.7, n, rng) for _ in range(3)])
        np.testing.assert_allclose(np_(bk.autocorr(x)), od.autocorr_batch(x), **TOL)
    x32 = torch.as_tensor(od.sample_ar1(0.5, 5000, rng), dtype=torch.float32, device="cuda")
    np.testing.assert_allclose(np_(bk.autocorr(x32)), od.autocorr(np_(x32).astype(np.float64)), **TOL)
