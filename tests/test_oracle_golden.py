"""Pin the oracle: its restatement must reproduce every committed fixture that
``oracle/gen_golden.py`` recorded from the UNMODIFIED reference (recorded RNG
streams in, reference trajectories / decisions / indices out)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import diagnostics as od
from oracle import samplers as osm
from oracle.models import build_model

RTOL = 1e-12  # fp64, identical op order: observed 0.0


def _names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def _close(a, b):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-14)


@pytest.mark.parametrize("name", _names("hmc_"))
def test_hmc(name):
    z = golden(name)
    metric = None if z["metric"].size == 0 else z["metric"]
    d, l, a = osm.hmc_diag_batch(build_model(z), z["theta0"], z["normals"], z["uniforms"],
                                 float(z["stepsize"]), int(z["steps"]), metric)
    _close(d, z["draws"]); _close(l, z["logps"])
    assert np.array_equal(a, z["accepts"])


@pytest.mark.parametrize("name", _names("mala_"))
def test_mala(name):
    z = golden(name)
    d, l, a = osm.mala_batch(build_model(z), z["theta0"], z["normals"], z["uniforms"],
                             float(z["epsilon"]))
    _close(d, z["draws"]); _close(l, z["logps"])
    assert np.array_equal(a, z["accepts"])


@pytest.mark.parametrize("name", _names("metropolis_") + _names("mh_"))
def test_metropolis(name):
    z = golden(name)
    d, l, a = osm.metropolis_rw_batch(build_model(z), z["theta0"], z["normals"], z["uniforms"],
                                      float(z["scale"]), bool(z["hastings"]))
    _close(d, z["draws"]); _close(l, z["logps"])
    assert np.array_equal(a, z["accepts"])


@pytest.mark.parametrize("name", _names("drghmc_"))
def test_drghmc(name):
    z = golden(name)
    d, l, r, u = osm.drghmc_batch(build_model(z), z["theta0"], z["rho0"], z["normals"],
                                  np.nan_to_num(z["uniforms"], nan=0.5), int(z["max_proposals"]),
                                  z["step_sizes"].tolist(), z["step_counts"].tolist(),
                                  float(z["damping"]), None, bool(z["prob_retry"]))
    _close(d, z["draws"]); _close(l, z["logps"]); _close(r, z["rho_final"])
    assert np.array_equal(u, z["n_used"])


@pytest.mark.parametrize("name", _names("smc_"))
def test_smc(name):
    z = golden(name)
    kernel = None
    if "kernel_kind" in z.files:      # MALA / HMC kernels inside the SMC (recorded with the reference's own classes)
        kind = str(z["kernel_kind"])
        kernel = (kind, float(z["scale"])) if kind == "mala" else (kind, float(z["scale"]), int(z["kernel_steps"]))
    th, idx = osm.smc_tempered(build_model(z), z["thetas0"], z["normals"], z["acc_uniforms"],
                               z["res_uniforms"], float(z["scale"]), int(z["T"]), kernel=kernel)
    assert np.array_equal(idx, z["indices"])
    _close(th, z["thetas_final"])


@pytest.mark.parametrize("tag", ["n37", "n1000", "n10000"])
def test_diagnostics(tag):
    z = golden("diagnostics")
    x = z["x_" + tag]
    _close(od.autocorr_batch(x), z["autocorr_" + tag])
    _close(od.iat_ess_batch(x, "ipse")[0], z["iat_ipse_" + tag])
    t, e = od.iat_ess_batch(x, "imse")
    _close(t, z["iat_imse_" + tag]); _close(e, z["ess_" + tag])
    _close(od.iat_ess_batch(x, "ipse")[1], z["ess_ipse_" + tag])


def test_rhat():
    z = golden("diagnostics")
    _close(od.rhat_batch(z["rhat_chains"]), z["rhat"])
    ch = z["rhat_chains"]
    _close([od.split_rhat(list(ch[:, :, p])) for p in range(3)], z["split_rhat"])


def test_rank_normalized_rhat_golden():
    """Oracle restatement vs the live reference's recorded outputs (rhat.py:27-108, 205-236)."""
    z = golden("diagnostics")
    ch = z["rhat_chains"]
    for p in range(3):
        chains = list(ch[:, :, p])
        assert np.array_equal(np.stack(od.rank_chains(chains)), z["ranks"][:, :, p])
        _close(np.stack(od.rank_normalize_chains(chains)), z["rank_normalized"][:, :, p])
        _close(od.rank_normalized_rhat(chains), z["rank_normalized_rhat"][p])
    heavy = list(z["cauchy_chains"])
    _close(od.split_rhat(heavy), z["cauchy_split_rhat"])
    _close(np.stack(od.rank_normalize_chains(heavy)), z["cauchy_rank_normalized"])
    _close(od.rank_normalized_rhat(heavy), z["cauchy_rank_normalized_rhat"])
