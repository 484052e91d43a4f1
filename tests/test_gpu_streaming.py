"""the sampler -> diagnostics hand-off without a second pass or a transposition.
 * streaming moments folded INSIDE the fused sampling kernel (registers across the draws of a launch, one Chan
   merge per launch): equal to the moments of the stored draws, draws unchanged, one kernel launch per call;
 * series-major draw output [C, D, n] written by the sampler (shared-memory staging of 8 draws per series):
   equal to the transposed [n, C, D] output, and consumed in place by ess / autocorr."""
import numpy as np
import pytest
import torch

from _dev import np_

pytestmark = pytest.mark.gpu


def _mk(bk, kind, C, seed):
    rng = np.random.default_rng(3)
    if kind == "hmc_iso":                  # c1's kernel: 8 lanes x 16 elements
        return bk.HMCDiag(bk.IsoGauss(100), 0.1, 10, chains=C, seed=seed)
    if kind == "hmc_diag_offset":          # means far from zero: the shifted sums must not cancel
        D = 37
        return bk.HMCDiag(bk.DiagGauss(rng.normal(size=D) * 3 + 1000.0, rng.uniform(0.5, 3, D)), 0.2, 5, chains=C, seed=seed)
    if kind == "mala_iso":
        return bk.MALA(bk.IsoGauss(50), 0.05, chains=C, seed=seed)
    if kind == "mh_small":                 # D <= 4: one lane per chain
        return bk.Metropolis(bk.IsoGauss(3), bk.GaussianRW(0.5), chains=C, seed=seed)
    return bk.HMCDiag(bk.DiagGauss(np.zeros(300), rng.uniform(0.5, 3, 300)), 0.1, 4, chains=C, seed=seed)   # 32 x 2 layout


@pytest.mark.parametrize("kind", ["hmc_iso", "hmc_diag_offset", "mala_iso", "mh_small", "hmc_d300"])
def test_fused_moments_equal_moments_of_the_draws(bk, kind):
    C = 333
    a, b = _mk(bk, kind, C, 4), _mk(bk, kind, C, 4)
    assert a._fuses_extras()
    kept = []
    lib = bk._lib.lib()
    for n in (7, 1, 16, 40):
        d, lp = a.sample_n(n, moments=True)
        d2, lp2 = b.sample_n(n)                                  # same seed, moments off: the chains must not change
        assert torch.equal(d, d2) and torch.equal(lp, lp2)
        kept.append(d)
    l0 = lib.bk_launch_count()
    a.sample_n(25, keep_draws=False, moments=True)               # monitored, no draw buffer at all
    assert lib.bk_launch_count() - l0 == 1                       # the sampling kernel IS the moments pass
    d2, _ = b.sample_n(25)
    kept.append(d2)
    x = torch.cat(kept).double()                                 # all 89 monitored draws
    mean, var, n = a.running_moments()
    assert n == 89
    np.testing.assert_allclose(np_(mean), np_(x.mean(0)), rtol=2e-7, atol=1e-6)
    np.testing.assert_allclose(np_(var), np_(x.var(0, unbiased=True)), rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(np_(a.running_rhat()), np_(bk.rhat(torch.cat(kept), draws_first=True)), rtol=1e-4)


@pytest.mark.parametrize("kind", ["hmc_iso", "mala_iso", "mh_small", "hmc_d300"])
@pytest.mark.parametrize("n", [1, 8, 13, 64])
def test_series_major_draws(bk, kind, n):
    C = 150
    a, b = _mk(bk, kind, C, 9), _mk(bk, kind, C, 9)
    for _ in range(2):                                           # second call: state carried on
        ds, lps = a.sample_n(n, layout="series", moments=True)
        d, lp = b.sample_n(n)
        assert ds.shape == (C, d.shape[2], n)
        assert torch.equal(ds.permute(2, 0, 1), d) and torch.equal(lps, lp)
    assert torch.equal(a.theta, b.theta)
    if n >= 8:
        D = d.shape[2]
        e_series = bk.ess(ds.reshape(C * D, n)).reshape(C, D)   # contiguous series: streamed in place, no gather
        e_draws = bk.ess(d, draws_first=True)
        np.testing.assert_allclose(np_(e_series), np_(e_draws), rtol=1e-9)
        k = min(5, D)
        ac = bk.autocorr(ds.reshape(C * D, n)[:k])
        want = bk.autocorr(d[:, 0, :k].t().contiguous())
        np.testing.assert_allclose(np_(ac), np_(want), rtol=0, atol=1e-12)


def test_series_layout_needs_the_fused_engine(bk):
    from oracle.models import DensePrecGauss
    s = bk.HMCDiag(bk.DensePrecGauss(DensePrecGauss.c2_precision(128, 1)), 0.1, 3, chains=64, seed=0)
    assert not s._fuses_extras()
    with pytest.raises(NotImplementedError):
        s.sample_n(4, layout="series")
    s.sample_n(4, moments=True)                                   # GEMM engine: folds the draws it wrote
    assert s.running_moments()[2] == 4
    with pytest.raises(ValueError):
        s.sample_n(4, layout="bogus")
