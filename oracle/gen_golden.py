"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the authoring container
(the reference lives at /root/reference there):

    python -m oracle.gen_golden

For every fixture the reference sampler runs under a recording RNG proxy
(oracle/record.py); the recorded streams are then fed to the oracle
restatement (oracle/samplers.py, oracle/diagnostics.py) and the two are
required to agree (trajectories to 1e-12 relative, accept decisions and
resample indices exactly) before the fixture is written.  Fixtures hold only
derived data (inputs, recorded streams, reference outputs) -- no reference
source.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import diagnostics as od
from . import samplers as osm
from .models import (Binomial, DensePrecGauss, DiagGauss, GaussPriorLik, HierLogReg, IsoGauss,
                     Tempered, model_spec)
from .record import LegacyRecorder, RecordingGenerator, load_reference, record_chain

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stack_runs(runs):
    draws = np.stack([r["draws"] for r in runs], 1)       # [n, C, D]
    logps = np.stack([r["logps"] for r in runs], 1)       # [n, C]
    return draws, logps


def _check(name, got, want, rtol=1e-12):
    got, want = np.asarray(got), np.asarray(want)
    scale = np.maximum(np.abs(want), 1.0)
    err = float(np.max(np.abs(got - want) / scale)) if want.size else 0.0
    print(f"  {name:34s} max rel err {err:.2e}")
    assert err <= rtol, (name, err)


def gen_hmc(bk, tag, model, n, C, eps, L, metric=None, D=None):
    D = model.dims()
    runs, th0 = [], []
    for c in range(C):
        def mk():
            s = bk.HMCDiag(model, eps, L, metric_diag=metric, seed=1000 + c)
            th0.append(np.array(s._theta, copy=True))
            return s
        runs.append(record_chain(mk, n, seed=c))
    draws, logps = _stack_runs(runs)
    normals = np.stack([np.stack([r["normals"][t][0] for t in range(n)]) for r in runs], 1)
    uniforms = np.stack([np.array([r["uniforms"][t][0] for t in range(n)]) for r in runs], 1)
    theta0 = np.stack(th0)
    o_d, o_l, o_a = osm.hmc_diag_batch(model, theta0, normals, uniforms, eps, L, metric)
    print(tag)
    _check("draws", o_d, draws)
    _check("joint logp", o_l, logps)
    ref_acc = np.stack([r["accepts"] for r in runs], 1)
    if L == 0:  # leapfrog returns the same array object: accepts unobservable
        ref_acc = o_a
    assert np.array_equal(ref_acc, o_a), "accept decisions differ"
    print(f"  accept rate {ref_acc.mean():.3f}")
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), theta0=theta0, normals=normals,
                        uniforms=uniforms, draws=draws, logps=logps, accepts=ref_acc,
                        stepsize=eps, steps=L,
                        metric=np.zeros(0) if metric is None else np.asarray(metric))


def gen_mala(bk, tag, model, n, C, eps):
    runs, th0 = [], []
    for c in range(C):
        def mk():
            s = bk.MALA(model, eps, seed=2000 + c)
            th0.append(np.array(s._theta, copy=True))
            return s
        runs.append(record_chain(mk, n, seed=50 + c))
    draws, logps = _stack_runs(runs)
    normals = np.stack([np.stack([r["normals"][t][0] for t in range(n)]) for r in runs], 1)
    uniforms = np.stack([np.array([r["uniforms"][t][0] for t in range(n)]) for r in runs], 1)
    theta0 = np.stack(th0)
    o_d, o_l, o_a = osm.mala_batch(model, theta0, normals, uniforms, eps)
    print(tag)
    _check("draws", o_d, draws)
    _check("logp", o_l, logps)
    ref_acc = np.stack([r["accepts"] for r in runs], 1)
    assert np.array_equal(ref_acc, o_a)
    print(f"  accept rate {ref_acc.mean():.3f}")
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), theta0=theta0, normals=normals,
                        uniforms=uniforms, draws=draws, logps=logps, accepts=ref_acc,
                        epsilon=eps)


def gen_metropolis(bk, tag, model, n, C, scale, hastings):
    runs, th0, props = [], [], []
    for c in range(C):
        pg = RecordingGenerator(7000 + c)
        props.append(pg)

        def mk():
            fn = lambda th: pg.normal(loc=th, scale=scale)
            if hastings:
                tl = lambda to, frm: -0.5 * np.dot(to - frm, to - frm) / (scale * scale)
                s = bk.MetropolisHastings(model, fn, tl, seed=3000 + c)
            else:
                s = bk.Metropolis(model, fn, seed=3000 + c)
            th0.append(np.array(s._theta, copy=True))
            return s
        runs.append(record_chain(mk, n, seed=90 + c))
    draws, logps = _stack_runs(runs)
    normals = np.stack([np.stack(pg.normals) for pg in props], 1)
    uniforms = np.stack([np.array([r["uniforms"][t][0] for t in range(n)]) for r in runs], 1)
    theta0 = np.stack(th0)
    o_d, o_l, o_a = osm.metropolis_rw_batch(model, theta0, normals, uniforms, scale, hastings)
    print(tag)
    _check("draws", o_d, draws)
    _check("logp", o_l, logps)
    ref_acc = np.stack([r["accepts"] for r in runs], 1)
    assert np.array_equal(ref_acc, o_a)
    print(f"  accept rate {ref_acc.mean():.3f}")
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), theta0=theta0, normals=normals,
                        uniforms=uniforms, draws=draws, logps=logps, accepts=ref_acc,
                        scale=scale, hastings=hastings)


def gen_drghmc(bk, tag, model, n, C, K, sizes, counts, damping, prob_retry):
    runs, th0, rh0 = [], [], []
    for c in range(C):
        def mk():
            s = bk.DrGhmcDiag(model, K, list(sizes), list(counts), damping,
                              seed=4000 + c, prob_retry=prob_retry)
            th0.append(np.array(s._theta, copy=True))
            rh0.append(np.array(s._rho, copy=True))
            return s
        runs.append(record_chain(mk, n, seed=130 + c))
    draws, logps = _stack_runs(runs)
    D = model.dims()
    normals = np.stack([np.stack([r["normals"][t][0] for t in range(n)]) for r in runs], 1)
    uniforms = np.full((n, C, 2 * K), np.nan)
    n_used = np.zeros((n, C), dtype=np.int64)
    for c, r in enumerate(runs):
        for t in range(n):
            u = r["uniforms"][t]
            uniforms[t, c, :len(u)] = u
            n_used[t, c] = len(u)
    theta0, rho0 = np.stack(th0), np.stack(rh0)
    rho_final = np.stack([r["sampler"]._rho for r in runs])
    o_d, o_l, o_r, o_u = osm.drghmc_batch(model, theta0, rho0, normals,
                                          np.nan_to_num(uniforms, nan=0.5), K, sizes,
                                          counts, damping, None, prob_retry)
    print(tag)
    _check("draws", o_d, draws)
    _check("joint logp", o_l, logps)
    _check("rho final", o_r, rho_final)
    assert np.array_equal(o_u, n_used), "uniform consumption differs"
    moved = np.stack([r["accepts"] for r in runs], 1)
    print(f"  move rate {moved.mean():.3f}; uniforms/draw hist "
          f"{np.bincount(n_used.ravel(), minlength=2 * K + 1).tolist()}")
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), theta0=theta0, rho0=rho0,
                        normals=normals, uniforms=uniforms, n_used=n_used, draws=draws,
                        logps=logps, rho_final=rho_final, accepts=moved, max_proposals=K,
                        step_sizes=np.asarray(sizes, dtype=np.float64),
                        step_counts=np.asarray(counts, dtype=np.int64), damping=damping,
                        prob_retry=prob_retry)


def gen_smc(bk, tag, model, M, T, scale):
    from bayes_kit.smc import metropolis_kernel
    D = model.dims()
    init_rng = np.random.default_rng(11)
    thetas0 = init_rng.normal(size=(M, D))
    np.random.seed(2024)
    with LegacyRecorder() as rec:
        smc = bk.TemperedLikelihoodSMC(model, M, T, lambda m: thetas0[m].copy(),
                                       metropolis_kernel(scale))
        smc.run()
    normals = np.stack(rec.normals).reshape(T, M, D)
    acc_u = np.array(rec.uniforms).reshape(T, M)
    res_u = np.stack(rec.res_uniforms)
    idx = np.stack(rec.indices)
    o_th, o_idx = osm.smc_tempered(model, thetas0, normals, acc_u, res_u, scale, T)
    print(tag)
    assert np.array_equal(o_idx, idx), "resample indices differ"
    assert np.array_equal(o_th, smc.thetas), "final particles differ"
    print(f"  indices identical ({idx.size}); unique final particles "
          f"{len(np.unique(smc.thetas[:, 0]))}/{M}")
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), thetas0=thetas0, normals=normals,
                        acc_uniforms=acc_u, res_uniforms=res_u, indices=idx,
                        thetas_final=np.asarray(smc.thetas), scale=scale, T=T)


class _LegacyRng:
    """`_rng` stand-in for a reference sampler used as an SMC kernel: draws from the process-global legacy
    stream like smc.py's own kernel does (smc.py:81,85), through the entry points LegacyRecorder patches."""

    def normal(self, loc=0.0, scale=1.0, size=None):
        return np.random.normal(loc=loc, scale=scale, size=size)

    def uniform(self):
        return np.random.uniform()


def gen_smc_kernel(bk, tag, model, M, T, kernel):
    """TemperedLikelihoodSMC (unmodified) with an MCMC kernel built from the reference's OWN MALA / HMCDiag
    classes on the tempered model -- the plug-in the TODO at smc.py:78 asks for.  kernel = ("mala", eps) or
    ("hmc", eps, L).  The kernel closure counts its calls to know the temperature (smc.py:54-57 hands it only
    theta and the log density)."""
    D = model.dims()
    init_rng = np.random.default_rng(13)
    thetas0 = init_rng.normal(size=(M, D))
    calls = [0]

    def kern(theta, lpminus1):
        n = 1 + calls[0] // M
        calls[0] += 1
        tm = Tempered(model, (n - 1) / T)
        assert abs(tm.log_density(theta) - lpminus1(theta)) <= 1e-12 * max(1.0, abs(lpminus1(theta)))
        if kernel[0] == "mala":
            s = bk.MALA(tm, kernel[1], init=np.array(theta, copy=True))
        else:
            s = bk.HMCDiag(tm, kernel[1], kernel[2], init=np.array(theta, copy=True))
        s._rng = _LegacyRng()
        return np.array(s.sample()[0], copy=True)

    np.random.seed(77)
    with LegacyRecorder() as rec:
        smc = bk.TemperedLikelihoodSMC(model, M, T, lambda m: thetas0[m].copy(), kern)
        smc.run()
    normals = np.stack(rec.normals).reshape(T, M, D)
    acc_u = np.array(rec.uniforms).reshape(T, M)
    res_u = np.stack(rec.res_uniforms)
    idx = np.stack(rec.indices)
    o_th, o_idx = osm.smc_tempered(model, thetas0, normals, acc_u, res_u, 0.0, T, kernel=kernel)
    print(tag)
    assert np.array_equal(o_idx, idx), "resample indices differ"
    _check("final particles", o_th, np.asarray(smc.thetas), 1e-13)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **model_spec(model), thetas0=thetas0, normals=normals,
                        acc_uniforms=acc_u, res_uniforms=res_u, indices=idx, thetas_final=np.asarray(smc.thetas),
                        scale=float(kernel[1]), T=T, kernel_kind=kernel[0],
                        kernel_steps=int(kernel[2]) if len(kernel) > 2 else 0)


def check_binomial_against_upstream():
    """oracle.models.Binomial == the reference's test/models/binomial.py densities (scipy), 1e-12."""
    sys.path.insert(0, "/root/reference")
    try:
        from test.models.binomial import Binomial as UpBinomial
    finally:
        sys.path.remove("/root/reference")
    up, mine = UpBinomial(alpha=2, beta=3, x=5, N=15), Binomial(2, 3, 5, 15)
    for th in np.linspace(-6, 6, 25):
        v = np.array([th])
        assert abs(up.log_prior(v) - mine.log_prior(v)) < 1e-12 * max(1, abs(up.log_prior(v)))
        assert abs(up.log_likelihood(v) - mine.log_likelihood(v)) < 1e-12 * max(1, abs(up.log_likelihood(v)))
        lp, g = mine.log_density_gradient(v)
        fd = (mine.log_density(v + 1e-6) - mine.log_density(v - 1e-6)) / 2e-6
        assert abs(g[0] - fd) < 1e-6
    print("oracle Binomial == upstream test model (log_prior, log_likelihood), analytic gradient == FD")


def gen_round2(bk):
    """Fixtures added in round 2: DrGhmcDiag on the non-separable plugins (lockstep engine), the reference's
    own beta-binomial test target under SMC / DrGHMC / HMC, and MALA / HMC kernels inside the SMC."""
    rng = np.random.default_rng(2)
    check_binomial_against_upstream()
    P64 = DensePrecGauss.c2_precision(64)
    gen_drghmc(bk, "drghmc_dense_d64", DensePrecGauss(P64), n=14, C=4, K=3, sizes=[0.45, 0.2, 0.1], counts=[3, 5, 8],
               damping=0.4, prob_retry=True)
    gen_drghmc(bk, "drghmc_dense_d64_retry0", DensePrecGauss(P64, mu=rng.normal(size=64)), n=10, C=3, K=2,
               sizes=[0.5, 0.2], counts=[4, 6], damping=1.0, prob_retry=False)
    gen_drghmc(bk, "drghmc_dense_d1000", DensePrecGauss(DensePrecGauss.c2_precision(1000)), n=3, C=2, K=2,
               sizes=[0.12, 0.05], counts=[4, 8], damping=0.5, prob_retry=True)
    X, y = HierLogReg.c3_data(400, 6, seed=0)
    gen_drghmc(bk, "drghmc_hlr_n400_d6", HierLogReg(X, y), n=12, C=3, K=2, sizes=[0.12, 0.05], counts=[3, 6],
               damping=0.3, prob_retry=True)
    bn = Binomial(2, 3, 5, 15)
    gen_drghmc(bk, "drghmc_binom", bn, n=40, C=4, K=3, sizes=[0.9, 0.4, 0.1], counts=[3, 3, 3], damping=0.2,
               prob_retry=True)       # test_drghmc.py:151-176's target (larger first step: exercises the retries)
    gen_hmc(bk, "hmc_binom", bn, n=30, C=4, eps=0.2, L=4)               # test_hmc.py:68-86's target
    gen_mala(bk, "mala_binom", bn, n=30, C=4, eps=0.3)                  # test_mala.py:25-41's target
    gen_smc(bk, "smc_binom", bn, M=75, T=15, scale=0.5)                 # test_tempered_smc.py:8-30, literally
    mu = rng.normal(size=5)
    gpl = GaussPriorLik(np.zeros(5), np.ones(5), mu, 4.0 * np.ones(5))
    gen_smc_kernel(bk, "smc_gauss_d5_mala", gpl, M=60, T=5, kernel=("mala", 0.05))
    gen_smc_kernel(bk, "smc_gauss_d5_hmc", gpl, M=60, T=5, kernel=("hmc", 0.2, 3))
    gen_smc_kernel(bk, "smc_binom_hmc", bn, M=50, T=6, kernel=("hmc", 0.3, 2))


def gen_diagnostics(bk):
    from bayes_kit.autocorr import autocorr as r_autocorr
    from bayes_kit.rhat import rhat as r_rhat, split_rhat as r_split_rhat
    rng = np.random.default_rng(5)
    phis = np.array([0.0, 0.3, 0.6, 0.9, -0.5, -0.3, 0.1, 0.75])
    out = {}
    for N, tag in [(1000, "n1000"), (10000, "n10000"), (37, "n37")]:
        ph = phis if N != 10000 else phis[[0, 3]]
        x = np.stack([od.sample_ar1(p, N, rng) for p in ph])
        ac = np.stack([r_autocorr(r) for r in x])
        ipse = np.array([bk.iat_ipse(r) for r in x])
        imse = np.array([bk.iat_imse(r) for r in x])
        ess = np.array([bk.ess(r) for r in x])
        ess_ip = np.array([bk.ess_ipse(r) for r in x])
        print(f"diagnostics {tag}")
        _check("autocorr", od.autocorr_batch(x), ac, 1e-13)
        _check("iat_ipse", od.iat_ess_batch(x, "ipse")[0], ipse, 1e-13)
        _check("iat_imse", od.iat_ess_batch(x, "imse")[0], imse, 1e-13)
        _check("ess", od.iat_ess_batch(x, "imse")[1], ess, 1e-13)
        out.update({f"x_{tag}": x, f"phi_{tag}": ph, f"autocorr_{tag}": ac,
                    f"iat_ipse_{tag}": ipse, f"iat_imse_{tag}": imse,
                    f"ess_{tag}": ess, f"ess_ipse_{tag}": ess_ip})
    ch = rng.normal(size=(8, 500, 3)) + np.array([0.0, 0.5, 2.0]) * rng.normal(size=(8, 1, 3))
    rh = np.array([r_rhat(list(ch[:, :, p])) for p in range(3)])
    srh = np.array([r_split_rhat(list(ch[:, :, p])) for p in range(3)])
    _check("rhat", od.rhat_batch(ch), rh, 1e-13)
    _check("split_rhat", [od.split_rhat(list(ch[:, :, p])) for p in range(3)], srh, 1e-13)
    out.update(rhat_chains=ch, rhat=rh, split_rhat=srh)
    # rank-normalised family (rhat.py:27-108, 205-236): continuous draws -> no ties
    from bayes_kit.rhat import (rank_chains as r_rank, rank_normalize_chains as r_rn,
                                rank_normalized_rhat as r_rnr)
    rk = np.stack([np.stack(r_rank(list(ch[:, :, p]))) for p in range(3)], -1)          # [8, 500, 3]
    rn = np.stack([np.array(r_rn(list(ch[:, :, p]))) for p in range(3)], -1)
    rnr = np.array([r_rnr(list(ch[:, :, p])) for p in range(3)])
    _check("rank_chains", np.stack([np.stack(od.rank_chains(list(ch[:, :, p]))) for p in range(3)], -1), rk, 0)
    _check("rank_normalize", np.stack([np.stack(od.rank_normalize_chains(list(ch[:, :, p]))) for p in range(3)], -1),
           rn, 1e-14)
    _check("rank_normalized_rhat", [od.rank_normalized_rhat(list(ch[:, :, p])) for p in range(3)], rnr, 1e-13)
    heavy = rng.standard_cauchy(size=(4, 101))            # odd length, heavy tails (the use case, rhat.py:211-214)
    heavy[2] += 3.0
    out.update(ranks=rk, rank_normalized=rn, rank_normalized_rhat=rnr, cauchy_chains=heavy,
               cauchy_split_rhat=r_split_rhat(list(heavy)), cauchy_rank_normalized_rhat=r_rnr(list(heavy)),
               cauchy_rank_normalized=np.array(r_rn(list(heavy))))
    _check("cauchy rank_normalized_rhat", od.rank_normalized_rhat(list(heavy)), r_rnr(list(heavy)), 1e-13)
    np.savez_compressed(os.path.join(OUT, "diagnostics.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    bk = load_reference()
    np.seterr(all="ignore")
    rng = np.random.default_rng(0)
    if "--round2" in sys.argv:          # only the fixtures added in round 2 (the others are unchanged)
        gen_round2(bk)
        return 0

    gen_hmc(bk, "hmc_iso_d100", IsoGauss(100), n=25, C=4, eps=0.1, L=10)
    gen_hmc(bk, "hmc_iso_d100_big_eps", IsoGauss(100), n=25, C=4, eps=0.8, L=7)
    gen_hmc(bk, "hmc_iso_d5_L0", IsoGauss(5), n=6, C=3, eps=0.3, L=0)
    gen_hmc(bk, "hmc_iso_d5_L1", IsoGauss(5, sigma=2.0), n=6, C=3, eps=0.3, L=1)
    gen_hmc(bk, "hmc_iso_d1_metric", IsoGauss(1), n=30, C=4, eps=0.4, L=5,
            metric=np.array([2.0]))
    gen_hmc(bk, "hmc_diag_d37", DiagGauss(rng.normal(size=37), rng.uniform(0.5, 4.0, 37)),
            n=20, C=5, eps=0.25, L=6)
    P64 = DensePrecGauss.c2_precision(64)
    gen_hmc(bk, "hmc_dense_d64", DensePrecGauss(P64), n=12, C=4, eps=0.2, L=8)
    gen_hmc(bk, "hmc_dense_d64_mu", DensePrecGauss(P64, mu=rng.normal(size=64)), n=8, C=3,
            eps=0.2, L=5)
    P1000 = DensePrecGauss.c2_precision(1000)
    gen_hmc(bk, "hmc_dense_d1000", DensePrecGauss(P1000), n=3, C=2, eps=0.05, L=10)

    gen_mala(bk, "mala_iso_d100", IsoGauss(100), n=30, C=4, eps=0.05)
    gen_mala(bk, "mala_diag_d37", DiagGauss(rng.normal(size=37), rng.uniform(0.5, 4.0, 37)),
             n=30, C=4, eps=0.05)
    gen_mala(bk, "mala_dense_d64", DensePrecGauss(P64), n=20, C=4, eps=0.08)
    gen_mala(bk, "mala_dense_d1000", DensePrecGauss(P1000), n=3, C=2, eps=1e-3)

    gen_metropolis(bk, "metropolis_iso_d10", IsoGauss(10), n=40, C=4, scale=0.3, hastings=False)
    gen_metropolis(bk, "mh_iso_d10", IsoGauss(10), n=40, C=4, scale=0.3, hastings=True)
    gen_metropolis(bk, "metropolis_dense_d64", DensePrecGauss(P64), n=20, C=3, scale=0.1,
                   hastings=False)

    for pr in (True, False):
        gen_drghmc(bk, f"drghmc_iso_d10_retry{int(pr)}", IsoGauss(10), n=40, C=4, K=3,
                   sizes=[1.7, 0.8, 0.4], counts=[2, 4, 8], damping=0.3, prob_retry=pr)
    gen_drghmc(bk, "drghmc_diag_d7_k4", DiagGauss(rng.normal(size=7), rng.uniform(0.5, 6.0, 7)),
               n=40, C=4, K=4, sizes=[1.2, 0.7, 0.35, 0.2], counts=[1, 2, 3, 5], damping=1.0,
               prob_retry=True)
    gen_drghmc(bk, "drghmc_iso_d100_k2", IsoGauss(100), n=15, C=3, K=2,
               sizes=[0.6, 0.2], counts=[5, 10], damping=0.5, prob_retry=True)

    mu = rng.normal(size=5)
    gen_smc(bk, "smc_gauss_d5", GaussPriorLik(np.zeros(5), np.ones(5), mu, 4.0 * np.ones(5)),
            M=96, T=6, scale=0.3)
    mu = rng.normal(size=50)
    gen_smc(bk, "smc_gauss_d50", GaussPriorLik(np.zeros(50), np.ones(50), mu, 4.0 * np.ones(50)),
            M=64, T=4, scale=0.1)

    gen_diagnostics(bk)

    # hierarchical logistic regression: no upstream definition -- pin our numpy
    # model by finite differences, then the reference HMCDiag drives it.
    X, y = HierLogReg.c3_data(400, 6, seed=0)
    hl = HierLogReg(X, y)
    th = rng.normal(size=8) * 0.3
    lp, g = hl.log_density_gradient(th)
    fd = np.array([(hl.log_density(th + 1e-6 * e) - hl.log_density(th - 1e-6 * e)) / 2e-6
                   for e in np.eye(8)])
    _check("hier_logreg grad vs FD", g, fd, 1e-6)
    gen_hmc(bk, "hmc_hlr_n400_d6", hl, n=15, C=3, eps=0.05, L=6)
    gen_round2(bk)
    print("all fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    sys.exit(main())
