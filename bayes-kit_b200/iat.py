"""Integrated autocorrelation time (reference: bayes_kit/iat.py)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._diag import as_series, call, finish, is_host


def _end_pos_pairs(acor) -> int:
    """Index one past the last initial positive pair (iat.py:7-43): the first
    even n with acor[n] + acor[n+1] < 0, else the largest even index reached.
    Host helper on a short sequence (the device kernels apply the same rule
    on the fly)."""
    n_pairs = len(acor) // 2
    for j in range(n_pairs):
        if acor[2 * j] + acor[2 * j + 1] < 0:
            return 2 * j
    return 2 * n_pairs


def _iat_ess(chain, estimator, device, draws_first, what):
    x, res, lay = as_series(chain, device, draws_first)
    if lay.n_draws < 4:
        raise ValueError(f"{what} requires len(chain) >= 4, but len(chain)={lay.n_draws}")
    t = torch.empty(lay.n_series, dtype=torch.float64, device=x.device)
    e = torch.empty_like(t)
    dt_ = L.BK_F32 if x.dtype == torch.float32 else L.BK_F64
    call(x, lambda lib, xp, dt, st, wp, wn: lib.bk_iat_ess(xp, dt, C.byref(lay), estimator, t.data_ptr(),
                                                          e.data_ptr(), wp, wn, st),
         ws_bytes=L.lib().bk_iat_ess_workspace_bytes(dt_, C.byref(lay)))
    h = is_host(chain)
    return finish(t, res, h), finish(e, res, h)


def iat_ipse(chain, device="cuda", draws_first=False):
    """Initial positive sequence estimator ``2 sum_{k<n} rho_k - 1`` (iat.py:46-92)."""
    return _iat_ess(chain, L.IAT_IPSE, device, draws_first, "iat")[0]


def iat_imse(chain, device="cuda", draws_first=False):
    """Initial monotone sequence estimator (iat.py:95-135)."""
    return _iat_ess(chain, L.IAT_IMSE, device, draws_first, "iat")[0]


def iat(chain, device="cuda", draws_first=False):
    """Delegates to iat_imse (iat.py:138-156)."""
    return iat_imse(chain, device, draws_first)
