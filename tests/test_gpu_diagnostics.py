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


def oracle_rhat(chains):
    return od.rhat([np.asarray(c, dtype=np.float64) for c in chains])


def od_rank(x):
    """[chains, draws] -> ranks in the same shape (oracle, stable ties)."""
    return np.stack(od.rank_chains(list(x)))


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


# ---- split / rank-normalised R-hat (reference rhat.py:9-108, 174-236) --------------
def test_split_and_rank_normalized_rhat_golden(bk):
    z = golden("diagnostics")
    ch = z["rhat_chains"]                                    # [8, 500, 3]
    for dt, tol in ((torch.float64, 1e-9), ):
        x = torch.as_tensor(ch, dtype=dt, device="cuda")
        np.testing.assert_allclose(np_(bk.split_rhat(x)), z["split_rhat"], rtol=tol)
        rk = np_(bk.rank_chains(x))
        assert np.array_equal(rk, z["ranks"])                # bit-exact integer work
        rn = np_(bk.rank_normalize_chains(x))
        np.testing.assert_allclose(rn, z["rank_normalized"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(np_(bk.rank_normalized_rhat(x)), z["rank_normalized_rhat"], rtol=tol)
        # the samplers' [draws, chains, params] layout, consumed in place
        xd = x.permute(1, 0, 2).contiguous()
        np.testing.assert_allclose(np_(bk.split_rhat(xd, draws_first=True)), z["split_rhat"], rtol=tol)
        assert np.array_equal(np_(bk.rank_chains(xd, draws_first=True)), z["ranks"])
        np.testing.assert_allclose(np_(bk.rank_normalized_rhat(xd, draws_first=True)), z["rank_normalized_rhat"],
                                   rtol=tol)
    heavy = z["cauchy_chains"]                               # [4, 101]: odd length, heavy tails
    assert abs(bk.split_rhat(list(heavy)) - float(z["cauchy_split_rhat"])) <= 1e-9 * float(z["cauchy_split_rhat"])
    np.testing.assert_allclose(np.stack(bk.rank_normalize_chains(list(heavy))), z["cauchy_rank_normalized"],
                               rtol=1e-12, atol=1e-13)
    got = bk.rank_normalized_rhat(list(heavy))
    assert abs(got - float(z["cauchy_rank_normalized_rhat"])) <= 1e-9


def test_rank_family_reference_known_answers(bk):
    """test_rhat.py:71-196 through the device path."""
    from scipy.stats import norm
    rn = lambda r, S: norm.ppf((r - 0.325) / (S - 0.25))
    assert bk.rank_chains([]) == [] and bk.rank_normalize_chains([]) == []
    for want, chains in [([[1]], [[2.3]]), ([[2, 3, 1]], [[3.9, 5.2, 2.1]]), ([[2], [1]], [[4.2], [1.9]]),
                         ([[2, 3], [5, 4], [1, 6]], [[4.2, 5.7], [7.2, 6.1], [-12.9, 107]]),
                         ([[2, 3, 6], [5, 4], [1]], [[4.2, 5.7, 108.0], [7.2, 6.1], [-12.9]])]:   # ragged
        for w, g in zip(want, bk.rank_chains(chains)):
            np.testing.assert_array_equal(w, g)
    np.testing.assert_allclose(bk.rank_normalize_chains([[32.7]])[0], [rn(1, 1)], rtol=1e-13)
    got = bk.rank_normalize_chains([[3.9, 3.1], [2.2, 5.9]])
    np.testing.assert_allclose(np.stack(got), [[rn(3, 4), rn(2, 4)], [rn(1, 4), rn(4, 4)]], rtol=1e-13)
    np.testing.assert_allclose(oracle_rhat([[1, 2], [3, 4]]), bk.split_rhat([[1, 2, 3, 4]]), rtol=1e-12)
    np.testing.assert_allclose(oracle_rhat([[1, -2, 3], [4, 5, 6], [7, 8], [9, 12]]),
                               bk.split_rhat([[1, -2, 3, 4, 5, 6], [7, 8, 9, 12]]), rtol=1e-12)
    r8 = [rn(i, 8) for i in range(1, 9)]
    np.testing.assert_allclose(bk.split_rhat([[r8[1], r8[2], r8[6], r8[7]], [r8[0], r8[3], r8[5], r8[4]]]),
                               bk.rank_normalized_rhat([[2, 3, 7, 8], [1, 4, 6, 5]]), rtol=1e-12)
    for fn, bad in [(bk.split_rhat, []), (bk.split_rhat, [[1, 2, 3]]), (bk.split_rhat, [[1, 2, 3, 4], [1, 2, 3]]),
                    (bk.rank_normalized_rhat, []), (bk.rank_normalized_rhat, [[1.01, 1.2, 1.3]]),
                    (bk.rank_normalized_rhat, [[1, 2, 3], [4]])]:
        with pytest.raises(ValueError):
            fn(bad)


def test_rank_sort_large_and_ties(bk):
    """Radix sort at a size that spans many tiles (fp32 and fp64 keys), negative values, and
    exact ties (rejected MCMC proposals repeat a draw): ties rank in flattened order."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((7, 9001)) * np.array([1e-3, 1, 1e3, 1, 1, 1e-20, 1])[:, None]
    x[3, 100:140] = x[3, 99]                                  # a run of rejections
    x[5, :10] = 0.0
    x[5, 3] = -0.0
    for dt in (torch.float64, torch.float32):
        xt = torch.as_tensor(x, dtype=dt, device="cuda")
        want = od_rank(np_(xt).astype(np.float64))
        assert np.array_equal(np_(bk.rank_chains(xt)), want)


@pytest.mark.parametrize("kind", ["hmc_iso", "hmc_dense", "mala_diag", "drghmc"])
def test_streaming_moments_match_rhat_on_stored_draws(bk, kind):
    """running per-chain moments folded batch by batch give the same R-hat
    (rhat.py:111-171) as the kernel that reads the stored [draws, chains, params] array."""
    from oracle.models import DensePrecGauss
    if kind == "hmc_iso":
        s = bk.HMCDiag(bk.IsoGauss(30), 0.3, 4, chains=64, seed=1)
    elif kind == "hmc_dense":
        s = bk.HMCDiag(bk.DensePrecGauss(DensePrecGauss.c2_precision(128, 1)), 0.1, 3, chains=300, seed=2)
    elif kind == "mala_diag":
        rng = np.random.default_rng(0)
        s = bk.MALA(bk.DiagGauss(rng.normal(size=17), rng.uniform(0.5, 2, 17)), 0.05, chains=33, seed=3)
    else:
        s = bk.DrGhmcDiag(bk.IsoGauss(12), 2, [0.4, 0.2], [3, 6], 0.5, chains=40, seed=4)
    kept = []
    for n in (5, 1, 23, 40):
        d, _ = s.sample_n(n, moments=True)
        kept.append(d)
    s.sample_n(7)                                              # not monitored
    _, lp = s.sample_n(20, keep_draws=False, moments=True)     # monitored, draws not returned
    assert lp.shape[0] == 20
    mean, var, n = s.running_moments()
    assert n == 5 + 1 + 23 + 40 + 20
    allx = torch.cat(kept)                                     # the first 69 monitored draws
    m_ref = allx.double().mean(0)
    # R-hat from moments of exactly the stored draws: re-accumulate them through the ABI in odd batches
    from bayes_kit_b200 import _lib as L
    C_, D = allx.shape[1], allx.shape[2]
    mu = torch.empty(C_, D, dtype=torch.float64, device="cuda"); m2 = torch.empty_like(mu)
    n0 = 0
    for a, b in ((0, 3), (3, 50), (50, 69)):
        chunk = allx[a:b].contiguous()
        L.check(L.lib().bk_moments_accumulate(chunk.data_ptr(), L.BK_F32, b - a, C_ * D, n0, mu.data_ptr(),
                                              m2.data_ptr(), torch.cuda.current_stream().cuda_stream))
        n0 += b - a
    np.testing.assert_allclose(np_(mu), np_(m_ref), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np_(m2 / 68), np_(allx.double().var(0, unbiased=True)), rtol=1e-10)
    from bayes_kit_b200.rhat import _rhat_from
    np.testing.assert_allclose(np_(_rhat_from(mu, m2 / 68, None, 69)), np_(bk.rhat(allx, draws_first=True)), **TOL)
    assert torch.isfinite(s.running_rhat()).all()
