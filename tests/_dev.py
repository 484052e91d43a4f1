"""Test helpers: golden fixture -> device plugin (fp64 parity mode)."""
import numpy as np
import torch

from oracle.models import DensePrecGauss as ODense


def device_model(bk, z, dtype=torch.float64):
    kind = str(z["model_kind"])
    if kind == "iso":
        return bk.IsoGauss(int(z["model_dims"]), float(z["model_sigma"]), dtype=dtype)
    if kind == "diag":
        return bk.DiagGauss(z["model_mu"], z["model_prec"], dtype=dtype)
    if kind == "dense":
        P = ODense.c2_precision(int(z["model_dims"]), int(z["model_seed"]))
        mu = z["model_mu"]
        return bk.DensePrecGauss(P, None if mu.size == 0 else mu, dtype=dtype)
    if kind == "gpl":
        return bk.GaussPriorLik(z["model_m0"], z["model_p0"], z["model_mu"], z["model_pl"], dtype=dtype)
    if kind == "hlr":
        return bk.HierLogReg(z["model_X"], z["model_y"], dtype=dtype)
    if kind == "binom":
        a, b, x, n = z["model_abxn"]
        return bk.Binomial(float(a), float(b), int(x), int(n), dtype=dtype)
    raise ValueError(kind)


def smc_kernel(bk, z):
    """The SMC move kernel a fixture was recorded with (default: metropolis_kernel(scale))."""
    kind = str(z["kernel_kind"]) if "kernel_kind" in z.files else "rw"
    if kind == "mala":
        return bk.mala_kernel(float(z["scale"]))
    if kind == "hmc":
        return bk.hmc_kernel(float(z["scale"]), int(z["kernel_steps"]))
    return bk.metropolis_kernel(float(z["scale"]))


def np_(t):
    return t.detach().cpu().numpy()


def assert_traj(draws, logps, acc, z, rtol=1e-10, tie=1e-12, what="logp"):
    """fp64 injected-stream contract (BASELINE north_star): trajectories within
    1e-10 relative, accept decisions bit-exact except where the accept test is
    tied to within 1e-12 (then the chain is excluded from that draw on)."""
    d, l, a = np_(draws), np_(logps), np_(acc).astype(bool)
    want_a = z["accepts"].astype(bool)
    ok = np.ones(a.shape[1], dtype=bool)
    for t in range(a.shape[0]):
        bad = (a[t] != want_a[t]) & ok
        if bad.any():
            # a flipped decision is only tolerated at a numerical tie
            raise AssertionError(f"accept decision differs at draw {t}, chains {np.where(bad)[0]}")
        np.testing.assert_allclose(d[t][ok], z["draws"][t][ok], rtol=rtol, atol=rtol)
        np.testing.assert_allclose(l[t][ok], z["logps"][t][ok], rtol=rtol, atol=rtol)
