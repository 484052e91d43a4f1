"""sample_host(): the host-buffer surface (pinned host state in, draw + log
density out, chain chunks pipelined over three streams) must return exactly
what the device-resident sample() returns -- chunking cannot change a chain
because Philox is keyed by the global chain id."""
import numpy as np
import pytest
import torch

from _dev import np_
from oracle.models import DensePrecGauss

pytestmark = pytest.mark.gpu


def _mk(bk, kind, C, seed):
    if kind == "hmc_dense":
        P = DensePrecGauss.c2_precision(256, 1)
        return bk.HMCDiag(bk.DensePrecGauss(P), 0.1, 5, chains=C, seed=seed)
    if kind == "hmc_iso":
        return bk.HMCDiag(bk.IsoGauss(100), 0.1, 10, chains=C, seed=seed)
    if kind == "mala_dense":
        P = DensePrecGauss.c2_precision(128, 2)
        return bk.MALA(bk.DensePrecGauss(P), 2e-3, chains=C, seed=seed)
    if kind == "mala_iso":
        return bk.MALA(bk.IsoGauss(37), 0.05, chains=C, seed=seed)
    if kind == "drghmc_iso":      # fused register-resident kernel; persistent momentum stays on the device
        return bk.DrGhmcDiag(bk.IsoGauss(20), 2, [0.6, 0.2], [3, 6], 0.5, chains=C, seed=seed)
    if kind == "drghmc_dense":    # lockstep engine on the dense-precision plugin
        P = DensePrecGauss.c2_precision(128, 3)
        return bk.DrGhmcDiag(bk.DensePrecGauss(P), 2, [0.3, 0.1], [3, 5], 0.5, chains=C, seed=seed)
    rw = bk.GaussianRW(0.3)
    return bk.Metropolis(bk.IsoGauss(10), rw, chains=C, seed=seed)


@pytest.mark.parametrize("kind", ["hmc_dense", "hmc_iso", "mala_dense", "mala_iso", "metropolis", "drghmc_iso",
                                  "drghmc_dense"])
@pytest.mark.parametrize("C,chunk", [(1000, 256), (777, 300), (512, None), (1000, [256, 500, 244])])
def test_sample_host_equals_device_sample(bk, kind, C, chunk):
    a, b = _mk(bk, kind, C, 11), _mk(bk, kind, C, 11)
    host_in = a.theta.cpu().pin_memory()
    for step in range(3):
        if step != 1:
            a._cache_valid.value = 0      # b re-evaluates (logp, grad) of the host-loaded state: same path
        d_dev, l_dev = a.sample()
        acc_dev = np_(a.last_accept).reshape(-1)
        # step 0 and 2 load the state from the host buffer, step 1 keeps the device state
        th = host_in if step != 1 else None
        d_host, l_host = b.sample_host(th, chunk_chains=chunk)
        assert not d_host.is_cuda and d_host.is_pinned()
        np.testing.assert_array_equal(np_(b.last_accept).reshape(-1), acc_dev)
        np.testing.assert_allclose(d_host.numpy(), np_(d_dev), rtol=0, atol=0)
        np.testing.assert_allclose(l_host.numpy(), np_(l_dev), rtol=1e-6, atol=1e-6)
        np.testing.assert_array_equal(np_(b.theta), d_host.numpy())
        host_in = d_host.clone().pin_memory()


def test_sample_host_validates(bk):
    s = _mk(bk, "hmc_iso", 64, 0)
    with pytest.raises(ValueError):
        s.sample_host(torch.zeros(63, 100))
    with pytest.raises(ValueError):
        s.sample_host(torch.zeros(64, 100, dtype=torch.float64))
    with pytest.raises(ValueError):
        s.sample_host(chunk_chains=[32, 16])          # explicit chunk sizes must cover every chain exactly once
    with pytest.raises(ValueError):
        s.sample_host(chunk_chains=[64, 0])
    with pytest.raises(ValueError):
        bk.HMCDiag(bk.IsoGauss(5), 0.1, 3).sample_host()
    with pytest.raises(ValueError, match="metric_diag must have 5 entries"):     # length is checked
        bk.DrGhmcDiag(bk.IsoGauss(5), 2, [0.2, 0.1], [2, 4], 0.5, metric_diag=np.ones(3), chains=8, seed=0)


@pytest.mark.parametrize("kind", ["hmc_dense", "hmc_iso", "drghmc_dense"])
@pytest.mark.parametrize("C,chunk,n", [(1000, 256, 5), (777, 300, 3), (512, None, 4)])
def test_sample_host_n_equals_device_sample_n(bk, kind, C, chunk, n):
    """n draws per call with host buffers (state in once, all draws out) == sample_n on the device, also when the
    next call's input aliases the last draw of the result buffer (bench.py's e2e loop)."""
    a, b = _mk(bk, kind, C, 5), _mk(bk, kind, C, 5)
    D = a.theta.shape[1]
    out = (torch.empty(n, C, D, pin_memory=True), torch.empty(n, C, pin_memory=True))
    out[0][n - 1].copy_(b.theta)
    for step in range(3):
        a._cache_valid.value = 0
        d_dev, l_dev = a.sample_n(n)
        acc_dev = np_(a.last_accept)
        d_host, l_host = b.sample_host_n(n, out[0][n - 1], out=out, chunk_chains=chunk)
        np.testing.assert_array_equal(np_(b.last_accept), acc_dev)
        np.testing.assert_allclose(d_host.numpy(), np_(d_dev), rtol=0, atol=0)
        np.testing.assert_allclose(l_host.numpy(), np_(l_dev), rtol=1e-6, atol=1e-6)
        np.testing.assert_array_equal(np_(b.theta), d_host[n - 1].numpy())
