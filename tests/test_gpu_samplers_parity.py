"""GPU parity: the CUDA samplers, fed the reference's RECORDED normals/uniforms
in fp64, must reproduce the unmodified reference's trajectories (golden
fixtures) within 1e-10 relative with identical accept decisions, and agree
with the oracle restatement on fresh seeded inputs."""
import glob
import os

import numpy as np
import pytest
import torch

from _dev import assert_traj, device_model, np_, smc_kernel
from conftest import GOLDEN, golden
from oracle import samplers as osm
from oracle.models import build_model

pytestmark = pytest.mark.gpu


def _names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def _engines():
    return ["fused", "generic"]


@pytest.fixture(params=["fused", "generic"])
def engine(request, monkeypatch):
    monkeypatch.setenv("BK_FORCE_GENERIC", "1" if request.param == "generic" else "0")
    return request.param


@pytest.mark.parametrize("name", [n for n in _names("hmc_") if "hlr" not in n])
def test_hmc_golden(bk, name, engine):
    z = golden(name)
    model = device_model(bk, z)
    metric = None if z["metric"].size == 0 else z["metric"]
    s = bk.HMCDiag(model, float(z["stepsize"]), int(z["steps"]), metric_diag=metric, init=z["theta0"])
    n = z["normals"].shape[0]
    draws, logp = s.sample_n(n, normals=z["normals"], uniforms=z["uniforms"])
    assert_traj(draws, logp, s.last_accept, z)
    np.testing.assert_allclose(np_(s.theta), z["draws"][-1], rtol=1e-10, atol=1e-10)


def test_hmc_golden_sample_by_sample(bk):
    """sample() one draw at a time == sample_n (state carried between calls)."""
    z = golden("hmc_dense_d64")
    model = device_model(bk, z)
    s = bk.HMCDiag(model, float(z["stepsize"]), int(z["steps"]), init=z["theta0"])
    for t in range(z["normals"].shape[0]):
        d, l = s.sample_n(1, normals=z["normals"][t:t + 1], uniforms=z["uniforms"][t:t + 1])
        np.testing.assert_allclose(np_(d[0]), z["draws"][t], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(np_(l[0]), z["logps"][t], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", _names("mala_"))
def test_mala_golden(bk, name, engine):
    z = golden(name)
    s = bk.MALA(device_model(bk, z), float(z["epsilon"]), init=z["theta0"])
    draws, logp = s.sample_n(z["normals"].shape[0], normals=z["normals"], uniforms=z["uniforms"])
    assert_traj(draws, logp, s.last_accept, z)


@pytest.mark.parametrize("name", _names("metropolis_") + _names("mh_"))
def test_metropolis_golden(bk, name, engine):
    z = golden(name)
    prop = bk.GaussianRW(float(z["scale"]))
    model = device_model(bk, z)
    if bool(z["hastings"]):
        s = bk.MetropolisHastings(model, prop, prop.transition_lp, init=z["theta0"])
    else:
        s = bk.Metropolis(model, prop, init=z["theta0"])
    draws, logp = s.sample_n(z["normals"].shape[0], normals=z["normals"], uniforms=z["uniforms"])
    assert_traj(draws, logp, s.last_accept, z)


@pytest.mark.parametrize("name", _names("drghmc_"))
def test_drghmc_golden(bk, name, engine):
    """Separable fixtures run through the fused register-resident kernel AND the lockstep engine
    (per-chain predication); dense / regression / binomial fixtures always use the lockstep engine."""
    z = golden(name)
    K = int(z["max_proposals"])
    s = bk.DrGhmcDiag(device_model(bk, z), K, [float(v) for v in z["step_sizes"]],
                      [int(v) for v in z["step_counts"]], float(z["damping"]), init=z["theta0"],
                      prob_retry=bool(z["prob_retry"]))
    s.set_momentum(z["rho0"])
    u = np.nan_to_num(z["uniforms"], nan=0.5)
    draws, logp = s.sample_n(z["normals"].shape[0], normals=z["normals"], uniforms=u)
    assert np.array_equal(np_(s.last_n_uniform), z["n_used"]), "uniform consumption differs"
    assert_traj(draws, logp, s.last_accept, z)
    np.testing.assert_allclose(np_(s.rho), z["rho_final"], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("C,D,L", [(33, 100, 10), (7, 3, 4), (5, 130, 3), (3, 300, 2), (130, 16, 5)])
def test_hmc_vs_oracle_fresh(bk, C, D, L, engine):
    """Sizes/shapes beyond the fixtures (ragged C, every lane-group config)."""
    from oracle.models import DiagGauss
    rng = np.random.default_rng(C * 1000 + D)
    mu, prec = rng.normal(size=D), rng.uniform(0.5, 3.0, D)
    th0, n = rng.normal(size=(C, D)), 4
    zs, us = rng.standard_normal((n, C, D)), rng.random((n, C))
    od, ol, oa = osm.hmc_diag_batch(DiagGauss(mu, prec), th0, zs, us, 0.2, L)
    s = bk.HMCDiag(bk.DiagGauss(mu, prec, dtype=torch.float64), 0.2, L, init=th0)
    d, l = s.sample_n(n, normals=zs, uniforms=us)
    np.testing.assert_allclose(np_(d), od, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np_(l), ol, rtol=1e-10, atol=1e-10)
    assert np.array_equal(np_(s.last_accept).astype(bool), oa)


def test_dense_vs_oracle_fresh(bk):
    from oracle.models import DensePrecGauss
    rng = np.random.default_rng(7)
    D, C, n = 200, 70, 3
    P = DensePrecGauss.c2_precision(D, 3)
    th0 = rng.normal(size=(C, D))
    zs, us = rng.standard_normal((n, C, D)), rng.random((n, C))
    od, ol, oa = osm.hmc_diag_batch(DensePrecGauss(P), th0, zs, us, 0.1, 5)
    s = bk.HMCDiag(bk.DensePrecGauss(P, dtype=torch.float64), 0.1, 5, init=th0)
    d, l = s.sample_n(n, normals=zs, uniforms=us)
    np.testing.assert_allclose(np_(d), od, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np_(l), ol, rtol=1e-10, atol=1e-10)
    assert np.array_equal(np_(s.last_accept).astype(bool), oa)
    od, ol, oa = osm.mala_batch(DensePrecGauss(P), th0, zs, us, 0.01)
    s = bk.MALA(bk.DensePrecGauss(P, dtype=torch.float64), 0.01, init=th0)
    d, l = s.sample_n(n, normals=zs, uniforms=us)
    np.testing.assert_allclose(np_(d), od, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np_(l), ol, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", _names("smc_"))
def test_smc_golden(bk, name):
    """Bit-exact multinomial resample indices and final particles under the
    recorded legacy-RNG streams (allowing index flips only where a uniform is
    within 1e-12 of a CDF edge -- none occur in the fixtures)."""
    z = golden(name)
    model = device_model(bk, z)
    M, T = z["thetas0"].shape[0], int(z["T"])
    smc = bk.TemperedLikelihoodSMC(model, M, T, z["thetas0"], smc_kernel(bk, z))
    for n in range(1, T + 1):
        smc.transition(n, normals=z["normals"][n - 1], acc_uniforms=z["acc_uniforms"][n - 1],
                       res_uniforms=z["res_uniforms"][n - 1])
        assert np.array_equal(np_(smc.last_indices), z["indices"][n - 1]), f"indices differ at n={n}"
    np.testing.assert_allclose(np_(smc.thetas), z["thetas_final"], rtol=1e-10, atol=1e-10)


def test_smc_systematic_vs_oracle(bk):
    from oracle.models import GaussPriorLik
    rng = np.random.default_rng(3)
    D, M, T, scale = 6, 257, 5, 0.25
    mu = rng.normal(size=D)
    om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = rng.normal(size=(M, D))
    zs, au = rng.standard_normal((T, M, D)), rng.random((T, M))
    ru = rng.random((T, M))
    oth, oidx = osm.smc_tempered(om, th0, zs, au, ru, scale, T, resample="systematic")
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64)
    smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(scale), resample="systematic")
    for n in range(1, T + 1):
        smc.transition(n, normals=zs[n - 1], acc_uniforms=au[n - 1], res_uniforms=ru[n - 1, :1])
        assert np.array_equal(np_(smc.last_indices), oidx[n - 1])
    np.testing.assert_allclose(np_(smc.thetas), oth, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("thr", [0.5, 0.9, 1.0])
def test_smc_adaptive_resampling_vs_oracle(bk, thr):
    """ESS-triggered resampling (parity unpinned upstream): the device decision, the carried
    log-weights and the particles follow the oracle restatement under injected streams."""
    from oracle.models import GaussPriorLik
    rng = np.random.default_rng(11)
    D, M, T, scale = 5, 300, 12, 0.3
    mu = rng.normal(size=D)
    om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = rng.normal(size=(M, D))
    zs, au, ru = rng.standard_normal((T, M, D)), rng.random((T, M)), rng.random((T, M))
    oth, ologw, oflags = osm.smc_tempered_adaptive(om, th0, zs, au, ru, scale, T, thr)
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64)
    smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(scale), resample="systematic",
                                   ess_threshold=thr)
    for n in range(1, T + 1):
        smc.transition(n, normals=zs[n - 1], acc_uniforms=au[n - 1], res_uniforms=ru[n - 1, :1])
    assert smc.resampled == list(oflags)
    if thr < 1.0:
        assert 0 < sum(oflags) < T           # the rule actually skips some temperatures
    np.testing.assert_allclose(np_(smc.thetas), oth, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np_(smc.log_weights), ologw, rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError):
        bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(scale), ess_threshold=1.5)


def test_model_protocol(bk):
    """dims / log_density / log_density_gradient of every plugin vs the numpy models."""
    from oracle import models as om
    rng = np.random.default_rng(0)
    D = 37
    mu, prec = rng.normal(size=D), rng.uniform(0.5, 2, D)
    P = om.DensePrecGauss.c2_precision(D, 1)
    pairs = [(bk.IsoGauss(D, 1.5, dtype=torch.float64), om.IsoGauss(D, 1.5)),
             (bk.DiagGauss(mu, prec, dtype=torch.float64), om.DiagGauss(mu, prec)),
             (bk.DensePrecGauss(P, mu, dtype=torch.float64), om.DensePrecGauss(P, mu)),
             (bk.GaussPriorLik(mu, prec, -mu, 2 * prec, dtype=torch.float64),
              om.GaussPriorLik(mu, prec, -mu, 2 * prec))]
    th = rng.normal(size=(9, D))
    for dm, nm in pairs:
        assert dm.dims() == nm.dims() == D
        lp, g = dm.log_density_gradient(th)
        want = [nm.log_density_gradient(t) for t in th]
        np.testing.assert_allclose(np_(lp), [w[0] for w in want], rtol=1e-12)
        np.testing.assert_allclose(np_(g), np.stack([w[1] for w in want]), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(np_(dm.log_density(th)), [w[0] for w in want], rtol=1e-12)
        lp1, g1 = dm.log_density_gradient(th[0])
        assert isinstance(lp1, float) and abs(lp1 - want[0][0]) <= 1e-12 * abs(want[0][0])
    gp, ng = pairs[3]
    np.testing.assert_allclose(np_(gp.log_prior(th)), [ng.log_prior(t) for t in th], rtol=1e-12)
    np.testing.assert_allclose(np_(gp.log_likelihood(th)), [ng.log_likelihood(t) for t in th], rtol=1e-12)


def test_hier_logreg_model_and_hmc(bk):
    """Builder-defined hierarchical logistic regression (no upstream definition):
    the device plugin vs the FD-checked numpy model, then the reference's own
    HMCDiag trajectory on it (golden fixture)."""
    from oracle.models import HierLogReg
    rng = np.random.default_rng(0)
    X, y = HierLogReg.c3_data(1000, 37, seed=1)
    om = HierLogReg(X, y)
    th = rng.normal(size=(70, 39)) * 0.4
    for dt, tol in ((torch.float64, 1e-11), (torch.float32, 2e-4)):
        dm = bk.HierLogReg(X, y, dtype=dt)
        assert dm.dims() == 39
        lp, g = dm.log_density_gradient(th)
        want = [om.log_density_gradient(t) for t in th]
        wl, wg = np.array([w[0] for w in want]), np.stack([w[1] for w in want])
        np.testing.assert_allclose(np_(lp), wl, rtol=tol, atol=tol * 10)
        assert np.abs(np_(g) - wg).max() <= tol * max(1.0, np.abs(wg).max())
    z = golden("hmc_hlr_n400_d6")
    s = bk.HMCDiag(device_model(bk, z), float(z["stepsize"]), int(z["steps"]), init=z["theta0"])
    draws, logp = s.sample_n(z["normals"].shape[0], normals=z["normals"], uniforms=z["uniforms"])
    assert_traj(draws, logp, s.last_accept, z)
    # MALA on the same plugin vs the oracle on fresh streams
    n, C = 5, 3
    th0 = rng.normal(size=(C, 8)) * 0.3
    zs, us = rng.standard_normal((n, C, 8)), rng.random((n, C))
    od, ol, oa = osm.mala_batch(build_model(z), th0, zs, us, 0.002)
    sm = bk.MALA(device_model(bk, z), 0.002, init=th0)
    d, l = sm.sample_n(n, normals=zs, uniforms=us)
    np.testing.assert_allclose(np_(d), od, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np_(l), ol, rtol=1e-10, atol=1e-9)
