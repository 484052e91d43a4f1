"""fp32 / device-Philox mode: statistical recovery (posterior moments within
4 MCSE, BASELINE north_star), reproducibility, seed sensitivity, and the
reference's cross-algorithm equivalences -- mirroring test_hmc.py:38-65,
test_mala.py:9-60, test_metropolis.py:106-254, test_equivalencies.py:12-32,
test_drghmc.py:97-148, test_tempered_smc.py:8-30."""
import numpy as np
import pytest
import torch

from _dev import np_

pytestmark = pytest.mark.gpu


def _check_moments(bk, draws, mean, var):
    """|mean_hat - mean| <= 4 MCSE and |var_hat - var| <= 4 MCSE (Gaussian
    target), each MCSE from the ESS of its own series (x for the mean,
    (x - mean)^2 for the variance) estimated by the device diagnostics."""
    d = np_(draws).reshape(-1, draws.shape[-1]).astype(np.float64)
    ess_x = _ess_total(bk, draws)
    mt = torch.as_tensor(mean, dtype=draws.dtype, device=draws.device)
    ess_x2 = _ess_total(bk, (draws - mt) ** 2)
    mcse_mean = np.sqrt(var / ess_x)
    mcse_var = var * np.sqrt(2.0 / ess_x2)
    assert np.all(np.abs(d.mean(0) - mean) <= 4 * mcse_mean + 1e-3), np.abs(d.mean(0) - mean).max()
    assert np.all(np.abs(d.var(0, ddof=1) - var) <= 4 * mcse_var + 1e-3), \
        (np.abs(d.var(0, ddof=1) - var) / mcse_var).max()


def _ess_total(bk, draws):
    """draws [n, C, D] (sampler layout): sum over chains of per-chain ESS, per
    parameter.  Anti-correlated chains (HMC near a half period) make the Geyer
    estimate exceed n or go negative (IAT <= 0, as in the reference); such
    chains are counted as n draws -- conservative for the MCSE."""
    e = np_(bk.ess(draws, draws_first=True))
    n = draws.shape[0]
    return np.where((e <= 0) | (e > n), n, e).sum(0)


def test_hmc_std_normal(bk):  # test_hmc.py:38-51
    model = bk.IsoGauss(100)
    # test_hmc.py uses eps=0.25, L=10 (eps*L = 2.5: draws anti-correlated at -0.8);
    # eps*L = 1.5 mixes x and x^2 fast, which a 4-MCSE check needs
    s = bk.HMCDiag(model, 0.15, 10, chains=512, seed=1)
    s.sample_n(50, keep_draws=False)
    draws, _ = s.sample_n(400)
    _check_moments(bk, draws, np.zeros(100), np.ones(100))
    assert 0.8 < float(s.last_accept.float().mean()) <= 1.0


def test_mala_std_normal(bk):  # test_mala.py:9-23
    model = bk.IsoGauss(10)
    s = bk.MALA(model, 0.3, chains=512, seed=2)
    s.sample_n(100, keep_draws=False)
    draws, _ = s.sample_n(1000)
    _check_moments(bk, draws, np.zeros(10), np.ones(10))


def test_metropolis_diag(bk):  # test_metropolis.py:106-123
    mu, prec = np.array([1.0, -2.0, 0.5]), np.array([1.0, 4.0, 0.25])
    model = bk.DiagGauss(mu, prec)
    s = bk.Metropolis(model, bk.GaussianRW(1.0), chains=1024, seed=3)
    s.sample_n(300, keep_draws=False)
    draws, _ = s.sample_n(2000)
    _check_moments(bk, draws, mu, 1 / prec)


def test_drghmc_std_normal(bk):  # test_drghmc.py:97-117
    model = bk.IsoGauss(5)
    s = bk.DrGhmcDiag(model, 2, [1.9, 0.5 / 5], [10, 20], 0.9, chains=512, seed=4)
    s.sample_n(100, keep_draws=False)
    draws, _ = s.sample_n(1000)
    _check_moments(bk, draws, np.zeros(5), np.ones(5))


def test_dense_hmc_moments(bk):
    from oracle.models import DensePrecGauss
    D = 32
    P = DensePrecGauss.c2_precision(D, 5)
    cov = np.linalg.inv(P)
    s = bk.HMCDiag(bk.DensePrecGauss(P), 0.125, 8, chains=1024, seed=5)  # eps*L = 1: no resonance
    s.sample_n(60, keep_draws=False)
    draws, _ = s.sample_n(300)
    _check_moments(bk, draws, np.zeros(D), np.diag(cov))


def test_reproducible_and_seed_sensitive(bk):  # test_hmc.py:54-65, test_mala.py:44-60
    model = bk.IsoGauss(7)
    init = np.random.default_rng(0).normal(size=(6, 7))
    for mk in (lambda sd: bk.HMCDiag(model, 0.25, 10, init=init, seed=sd),
               lambda sd: bk.MALA(model, 0.3, init=init, seed=sd),
               lambda sd: bk.Metropolis(model, bk.GaussianRW(0.5), init=init, seed=sd),
               lambda sd: bk.DrGhmcDiag(model, 2, [0.9, 0.3], [3, 6], 0.5, init=init, seed=sd)):
        a, b, c = mk(123), mk(123), mk(321)
        da = torch.stack([a.sample()[0] for _ in range(25)])
        db = torch.stack([next(b)[0] for _ in range(25)])     # next() == sample()
        dc = torch.stack([c.sample()[0] for _ in range(25)])
        assert torch.equal(da, db)
        assert not torch.equal(da, dc)
        # one call of 25 draws == 25 calls of one draw (Philox counter = draw index)
        dd, _ = mk(123).sample_n(25)
        assert torch.equal(da, dd)


def test_hmc_one_step_is_mala(bk):  # test_equivalencies.py:12-32
    model = bk.IsoGauss(1, dtype=torch.float64)
    init = np.array([0.2])
    eps = 0.02
    hmc = bk.HMCDiag(model, eps, 1, init=init, seed=123)
    mala = bk.MALA(model, 0.5 * eps ** 2, init=init, seed=123)
    d1 = np_(hmc.sample_n(50)[0]); d2 = np_(mala.sample_n(50)[0])
    np.testing.assert_array_almost_equal(d1, d2)
    assert len(np.unique(d1)) > 20


def test_metropolis_equals_mh_symmetric(bk):  # test_equivalencies.py:35-60
    model = bk.IsoGauss(3)
    init = np.zeros((4, 3))
    prop = bk.GaussianRW(0.7)
    a = bk.Metropolis(model, prop, init=init, seed=1848)
    b = bk.MetropolisHastings(model, prop, prop.transition_lp, init=init, seed=1848)
    assert torch.equal(a.sample_n(25)[0], b.sample_n(25)[0])


def test_single_chain_surface(bk):  # test_theta_initialization.py:34-54
    model = bk.IsoGauss(3)
    init = np.array([3.0, 3.0, 3.0])
    for s in (bk.HMCDiag(model, 0.25, 10, init=init), bk.MALA(model, 0.5, init=init),
              bk.Metropolis(model, bk.GaussianRW(1.0), init=init)):
        np.testing.assert_array_equal(np_(s.theta), init)
        th, lp = s.sample()
        assert th.shape == (3,) and lp.dim() == 0
    s = bk.HMCDiag(model, 0.25, 10, init=np.array([]))   # empty init == no init
    assert s.theta.shape == (3,)


def test_shard_invariance(bk):
    """Chains [lo, hi) of a sharded run == the same chains of the full run
    (Philox keyed by the global chain id)."""
    model = bk.IsoGauss(20)
    init = np.random.default_rng(1).normal(size=(64, 20))
    full = bk.HMCDiag(model, 0.2, 5, init=init, seed=9).sample_n(10)[0]
    part = bk.HMCDiag(model, 0.2, 5, init=init[40:64], seed=9, chain_offset=40).sample_n(10)[0]
    assert torch.equal(full[:, 40:64], part)
    # init=None: theta0 (and DrGhmcDiag's rho0) come from device Philox keyed by the global chain id too
    f = bk.HMCDiag(model, 0.2, 5, chains=64, seed=9)
    p = bk.HMCDiag(model, 0.2, 5, chains=24, seed=9, chain_offset=40)
    assert torch.equal(f.theta[40:64], p.theta)
    assert torch.equal(f.sample_n(3)[0][:, 40:64], p.sample_n(3)[0])
    assert abs(float(f.theta.mean())) < 0.2 and abs(float(f.theta.std()) - 1) < 0.1
    fd = bk.DrGhmcDiag(model, 2, [0.4, 0.2], [3, 6], 0.5, chains=64, seed=9)
    pd = bk.DrGhmcDiag(model, 2, [0.4, 0.2], [3, 6], 0.5, chains=24, seed=9, chain_offset=40)
    assert torch.equal(fd.rho[40:64], pd.rho) and torch.equal(fd.theta[40:64], pd.theta)
    assert not torch.equal(fd.rho, fd.theta)
    assert torch.equal(fd.sample_n(4)[0][:, 40:64], pd.sample_n(4)[0])


def test_smc_posterior(bk):  # test_tempered_smc.py:8-30 (Gaussian analogue)
    D, M, T = 2, 20000, 40
    mu = np.array([1.0, -0.5])
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = np.random.default_rng(0).normal(size=(M, D))
    for mode in ("multinomial", "systematic"):
        smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.5), resample=mode, seed=11)
        smc.run()
        th = np_(smc.thetas).astype(np.float64)
        np.testing.assert_allclose(th.mean(0), 0.8 * mu, atol=0.05)
        np.testing.assert_allclose(th.var(0, ddof=1), 0.2 * np.ones(D), atol=0.05)
        assert len(smc.weight_ess) == T and all(1 <= e <= M + 1e-6 for e in smc.weight_ess)


@pytest.mark.parametrize("D", [3, 12, 13, 100, 125, 200])
def test_engines_agree_in_philox_mode(bk, D, monkeypatch):
    """Device-RNG chains do not depend on which engine ran: the fused register-resident kernel and
    the generic row kernels follow one rule for the accept uniform (spare Philox block or the
    dedicated stream, common.cuh) and the same normal blocks."""
    res = []
    for force in ("0", "1"):
        monkeypatch.setenv("BK_FORCE_GENERIC", force)
        out = []
        for mk in (lambda: bk.HMCDiag(bk.IsoGauss(D, dtype=torch.float64), 0.2, 3, chains=70, seed=5),
                   lambda: bk.MALA(bk.IsoGauss(D, dtype=torch.float64), 0.05, chains=70, seed=6),
                   lambda: bk.Metropolis(bk.IsoGauss(D, dtype=torch.float64), bk.GaussianRW(0.3), chains=70, seed=7)):
            s = mk()
            d, _ = s.sample_n(6)
            out.append((np_(d), np_(s.last_accept)))
        res.append(out)
    for (d0, a0), (d1, a1) in zip(*res):
        assert np.array_equal(a0, a1)
        np.testing.assert_allclose(d0, d1, rtol=1e-12, atol=1e-12)


# ---- the reference's own beta-binomial test target (test/models/binomial.py) through every sampler --------------------
# Same model, sampler settings, burn-in and tolerances as upstream; device RNG, fp32; 256 chains in lockstep where the
# reference runs one (every chain must pass on average: the moments are pooled over chains AND draws, so the
# tolerances are far looser than needed -- they are the reference's).
def _binom(bk):
    return bk.Binomial(alpha=2, beta=3, x=5, N=15)


def _check_binom(model, draws, burn, atol_mean, atol_var):
    p = model.constrain_draws(draws[burn:]).double().reshape(-1)
    assert abs(float(p.mean()) - model.posterior_mean()) < atol_mean
    assert abs(float(p.var()) - model.posterior_variance()) < atol_var


def test_hmc_binom(bk):          # test_hmc.py:68-86
    model = _binom(bk)
    s = bk.HMCDiag(model, stepsize=0.08, steps=3, chains=256, seed=1)
    d, _ = s.sample_n(800)
    _check_binom(model, d, 100, 0.05, 0.01)


def test_mala_binom(bk):         # test_mala.py:25-41
    model = _binom(bk)
    s = bk.MALA(model, 0.07, chains=256, seed=2)
    d, _ = s.sample_n(1200)
    _check_binom(model, d, 200, 0.05, 0.01)


def test_metropolis_binom(bk):   # test_metropolis.py:127-139 (proposal normal(loc=theta, scale=4))
    model = _binom(bk)
    s = bk.Metropolis(model, bk.GaussianRW(4.0), chains=256, seed=3)
    d, _ = s.sample_n(1000)
    _check_binom(model, d, 0, 0.1, 0.1)


def test_drghmc_binom(bk):       # test_drghmc.py:151-176
    model = _binom(bk)
    s = bk.DrGhmcDiag(model, 3, [0.1] * 3, [3] * 3, 0.2, chains=256, seed=4)
    d, _ = s.sample_n(800)
    _check_binom(model, d, 100, 0.05, 0.01)


def test_rwm_smc_binom(bk):      # test_tempered_smc.py:8-30: M = 75, N = 15, metropolis_kernel(0.5); here 32 replicates pooled
    model = _binom(bk)
    g = torch.Generator(device="cuda").manual_seed(5)
    means, vars_ = [], []
    for rep in range(32):
        p0 = torch.distributions.Beta(2.0, 3.0).sample((75,)).to("cuda")          # initial_state: logit of a prior draw
        th0 = torch.log(p0 / (1 - p0)).reshape(75, 1)
        smc = bk.TemperedLikelihoodSMC(model, 75, 15, th0, bk.metropolis_kernel(0.5), seed=100 + rep)
        smc.run()
        p = model.constrain_draws(smc.thetas).double().reshape(-1)
        means.append(float(p.mean())); vars_.append(float(p.var()))
    assert abs(np.mean(means) - model.posterior_mean()) < 0.05        # the reference's tolerances for ONE run
    assert abs(np.mean(vars_) - model.posterior_variance()) < 0.01


# ---- lane layouts and leapfrog forms of the fused isotropic kernel (sampler_sep_kernel.cuh / sampler_sep_narrow.cu) -----
@pytest.mark.parametrize("D,L,eps", [(70, 9, 0.17), (84, 10, 0.15), (100, 7, 0.2), (108, 1, 0.3), (124, 10, 0.15),
                                     (50, 10, 0.15), (100, 2, 0.3)])
def test_hmc_iso_layouts_and_step_parity(bk, D, L, eps):
    """4 x J narrow layouts (D = 70 / 84 / 100 / 108: J = 5 / 6 / 7 / 7), the power-of-two ones (50, 124), odd and even
    L (the position-form recurrence ends in either of its two arrays) and L = 1: moments within 4 MCSE."""
    s = bk.HMCDiag(bk.IsoGauss(D), eps, L, chains=512, seed=11)
    s.sample_n(60, keep_draws=False)
    draws, _ = s.sample_n(300)
    _check_moments(bk, draws, np.zeros(D), np.ones(D))
    assert 0.7 < float(s.last_accept.float().mean()) <= 1.0


def test_hmc_iso_small_stepsize_keeps_the_force(bk):
    """eps^2 prec < 2^-10: the position-form coefficient 2 - eps^2 prec would round the force away in fp32, so the
    kernel takes the velocity form.  200 steps of eps = 0.01 must still move and sample N(0, 1)."""
    D = 100
    s = bk.HMCDiag(bk.IsoGauss(D), 0.01, 150, chains=512, seed=12)
    s.sample_n(20, keep_draws=False)
    draws, _ = s.sample_n(150)
    _check_moments(bk, draws, np.zeros(D), np.ones(D))
    assert float(s.last_accept.float().mean()) > 0.99


def test_hmc_iso_zero_steps_stays_put(bk):
    """hmc.py:43-44: with steps = 0 the two half kicks cancel and the proposal is the current point."""
    th0 = torch.randn(64, 100)
    s = bk.HMCDiag(bk.IsoGauss(100), 0.1, 0, init=th0, seed=13)
    d, _ = s.sample_n(3)
    assert torch.equal(d[-1].cpu(), th0.to(d.dtype))


@pytest.mark.parametrize("D", [70, 100, 108])
def test_mala_metropolis_narrow_layouts(bk, D):
    for s, n in ((bk.MALA(bk.IsoGauss(D), 0.05, chains=512, seed=14), 600),
                 (bk.Metropolis(bk.IsoGauss(D), bk.GaussianRW(0.25), chains=512, seed=15), 2500)):
        s.sample_n(n // 2, keep_draws=False)
        draws, _ = s.sample_n(n)
        _check_moments(bk, draws, np.zeros(D), np.ones(D))
