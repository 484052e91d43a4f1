"""Sharded TemperedLikelihoodSMC on ONE GPU ("fake world"): the G ranks of a multi-GPU run are
looped in lockstep in one process, through exactly the kernels / mailboxes / peer index stores / peer row
reads the NCCL-box run uses (peer pointers simply point at other ranks' buffers on the same device).
Proves: sharded == single-rank == oracle, bit for bit on indices (smc.py:60-75)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import samplers as osm
from _dev import device_model, np_

pytestmark = pytest.mark.gpu


def _run_fake(bk, world, make, T, streams):
    """make(group, lo, hi) -> smc for that rank; streams(n, lo, hi) -> (normals, acc_u, res_u) or Nones.
    Returns (indices [T, M], thetas [M, D], ranks)."""
    fw = bk.peer.FakeWorld(world)
    M = None
    ranks = []
    for r in range(world):
        s = make(fw.rank(r))
        ranks.append(s)
        M = s.M
    all_idx = []
    for n in range(1, T + 1):
        for s in ranks:
            z, au, _ = streams(n, s._lo, s._hi, s)
            s._move(n, z, au)
        for ph in (1, 2):
            for s in ranks:
                _, _, ru = streams(n, s._lo, s._hi, s)
                s._resample(n, ru, ph)
        torch.cuda.synchronize()
        all_idx.append(np.concatenate([np_(s.last_indices) for s in ranks]))
    th = np.concatenate([np_(s.thetas) for s in ranks])
    for s in ranks:
        s._check()
    return np.stack(all_idx), th, ranks


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("mode", ["systematic", "multinomial"])
def test_sharded_smc_equals_oracle(bk, world, mode):
    from oracle.models import GaussPriorLik
    rng = np.random.default_rng(5)
    D, M, T, scale = 6, 1031, 6, 0.25          # 1031 is prime: ragged shards for every world > 1
    mu = rng.normal(size=D)
    om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = rng.normal(size=(M, D))
    zs, au, ru = rng.standard_normal((T, M, D)), rng.random((T, M)), rng.random((T, M))
    oth, oidx = osm.smc_tempered(om, th0, zs, au, ru, scale, T, resample=mode)
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64)

    def make(group):
        return bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(scale), resample=mode, group=group)

    def streams(n, lo, hi, s):
        r = ru[n - 1, :1] if mode == "systematic" else ru[n - 1, lo:hi]
        return zs[n - 1, lo:hi], au[n - 1, lo:hi], r

    idx, th, ranks = _run_fake(bk, world, make, T, streams)
    assert np.array_equal(idx, oidx), f"indices differ (world={world}, {mode})"
    np.testing.assert_allclose(th, oth, rtol=1e-12, atol=1e-12)
    ess = ranks[0].weight_ess
    assert len(ess) == T and all(1 <= e <= M + 1e-6 for e in ess)
    for s in ranks[1:]:
        assert s.weight_ess == ess          # every rank derives the same global statistics


@pytest.mark.parametrize("name", ["smc_gauss_d5", "smc_gauss_d50"])
@pytest.mark.parametrize("world", [2, 5])
def test_sharded_smc_golden(bk, name, world):
    """The reference's own run (recorded legacy-RNG streams, tests/golden): multinomial indices bit-exact
    when the particles are sharded over 2 / 5 ranks."""
    z = golden(name)
    model = device_model(bk, z)
    M, T = z["thetas0"].shape[0], int(z["T"])

    def make(group):
        return bk.TemperedLikelihoodSMC(model, M, T, z["thetas0"], bk.metropolis_kernel(float(z["scale"])), group=group)

    def streams(n, lo, hi, s):
        return z["normals"][n - 1][lo:hi], z["acc_uniforms"][n - 1][lo:hi], z["res_uniforms"][n - 1][lo:hi]

    idx, th, _ = _run_fake(bk, world, make, T, streams)
    assert np.array_equal(idx, z["indices"])
    np.testing.assert_allclose(th, z["thetas_final"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_smc_adaptive(bk, world):
    from oracle.models import GaussPriorLik
    rng = np.random.default_rng(11)
    D, M, T, scale, thr = 5, 301, 12, 0.3, 0.5
    mu = rng.normal(size=D)
    om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = rng.normal(size=(M, D))
    zs, au, ru = rng.standard_normal((T, M, D)), rng.random((T, M)), rng.random((T, M))
    oth, ologw, oflags = osm.smc_tempered_adaptive(om, th0, zs, au, ru, scale, T, thr)
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64)

    def make(group):
        return bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(scale), resample="systematic",
                                        ess_threshold=thr, group=group)

    def streams(n, lo, hi, s):
        return zs[n - 1, lo:hi], au[n - 1, lo:hi], ru[n - 1, :1]

    _, th, ranks = _run_fake(bk, world, make, T, streams)
    assert ranks[0].resampled == list(oflags) and 0 < sum(oflags) < T
    np.testing.assert_allclose(th, oth, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.concatenate([np_(s.log_weights) for s in ranks]), ologw, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("world,D", [(2, 50), (8, 50), (3, 52), (2, 45)])
def test_sharded_smc_philox_invariant(bk, world, D):
    """fp32 device-RNG mode: the particles after T temperatures do not depend on the number of ranks
    (Philox is keyed by the global particle id, the fixed-point CDF is exact).  D = 50 / 52 / 45: the move kernel's
    64-bit, 128-bit (compile-time row access modes) and run-time row access paths."""
    M, T = 20000, 8
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = torch.randn(D, device="cuda", generator=g)
    th0 = torch.randn(M, D, device="cuda", generator=g)
    model = bk.GaussPriorLik(torch.zeros(D), torch.ones(D), mu, 4 * torch.ones(D))

    def make(group):
        return bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample="systematic", seed=7,
                                        group=group)

    none = lambda n, lo, hi, s: (None, None, None)
    idx1, th1, _ = _run_fake(bk, 1, make, T, none)
    idxg, thg, _ = _run_fake(bk, world, make, T, none)
    assert np.array_equal(idx1, idxg)
    assert np.array_equal(th1, thg)
    single = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample="systematic", seed=7)
    single.run()
    assert np.array_equal(np_(single.thetas), th1)
