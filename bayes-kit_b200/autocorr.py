"""autocorr (reference: bayes_kit/autocorr.py)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._diag import as_series, call


def autocorr(chain, device="cuda", draws_first=False):
    """Sample autocorrelation at lags 0..N-1 (biased estimator, ddof=0 variance;
    autocorr.py:6-33).  ``chain``: 1-D array-like (reference surface; returns a
    1-D float64 device tensor), or batched ``[chains, draws]`` /
    ``[chains, draws, params]`` (one series per (chain, param); returns
    ``[chains, N]`` / ``[chains, params, N]``).  Raises ValueError if there are
    fewer than 2 draws."""
    x, res, lay = as_series(chain, device, draws_first)
    N = lay.n_draws
    if N < 2:
        raise ValueError(f"autocorr requires len(chain) >= 2, but len(chain)={N}")
    out = torch.empty(lay.n_series, N, dtype=torch.float64, device=x.device)
    call(x, lambda lib, xp, dt, st, wp, wn: lib.bk_autocorr(xp, dt, C.byref(lay), out.data_ptr(), wp, wn, st),
         L.lib().bk_autocorr_workspace_bytes(lay.n_series, N))
    return out.reshape(res + (N,))
