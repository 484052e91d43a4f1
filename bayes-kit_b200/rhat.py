"""Potential scale reduction (reference: bayes_kit/rhat.py:111-171)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import dist as D_
from ._diag import as_series, call, is_host
from ._util import require_cuda, stream_ptr


def chain_moments(draws, device="cuda", draws_first=False):
    """Per-series (mean, ddof=1 variance) in fp64 -- rhat.py:165-166 batched."""
    x, res, lay = as_series(draws, device, draws_first)
    if lay.n_draws < 2:
        raise ValueError("rhat requires len(chain) >= 2 for every chain in chains")
    m = torch.empty(lay.n_series, dtype=torch.float64, device=x.device)
    v = torch.empty_like(m)
    call(x, lambda lib, xp, dt, st, wp, wn: lib.bk_chain_moments(xp, dt, C.byref(lay), m.data_ptr(),
                                                                v.data_ptr(), st))
    return m.reshape(res), v.reshape(res), lay.n_draws


def _rhat_from(mean, var, lengths, N):
    """mean/var [chains, params] f64 device; lengths [chains] int64 device or None."""
    n_chains, n_params = mean.shape
    if n_chains < 2:
        raise ValueError(f"rhat requires len(chains) >= 2, but len(chains) = {n_chains}")
    out = torch.empty(n_params, dtype=torch.float64, device=mean.device)
    with torch.cuda.device(mean.device):
        L.check(L.lib().bk_rhat_from_moments(mean.contiguous().data_ptr(), var.contiguous().data_ptr(),
                                             None if lengths is None else lengths.data_ptr(), int(N),
                                             n_chains, n_params, out.data_ptr(), stream_ptr(mean.device)))
    return out


def rhat(chains, device="cuda", draws_first=False, group=None):
    """R-hat = sqrt((nbar-1)/nbar + var(chain means, ddof=1) / mean(chain vars, ddof=1)).

    ``chains``: the reference's list of 1-D chains (ragged allowed, returns a
    float), a ``[chains, draws]`` array (float / 0-dim tensor) or
    ``[chains, draws, params]`` (tensor [params]).  With torch.distributed
    initialised and ``group`` given (or the default group), ``chains`` is this
    rank's shard of chains and the per-chain moments are all-gathered -- the
    only communication.  Raises ValueError for < 2 chains or a chain with < 2
    draws (rhat.py:157-162)."""
    host = is_host(chains)
    ragged = (isinstance(chains, (list, tuple)) and len(chains) > 0
              and len({len(c) for c in chains}) > 1)
    if isinstance(chains, (list, tuple)):
        if len(chains) < 2 and D_.rank_world(group)[1] == 1:
            raise ValueError(f"rhat requires len(chains) >= 2, but len(chains) = {len(chains)}")
        if not all(len(c) >= 2 for c in chains):
            raise ValueError("rhat requires len(chain) >= 2 for every chain in chains")
    if ragged:
        dev = require_cuda(device)
        ms, vs = [], []
        for c in chains:
            m, v, _ = chain_moments(np.asarray(c, dtype=np.float64), dev)
            ms.append(m.reshape(1)); vs.append(v.reshape(1))
        mean, var = torch.stack(ms), torch.stack(vs)
        lengths = torch.tensor([len(c) for c in chains], dtype=torch.int64, device=dev)
        return float(_rhat_from(mean, var, lengths, 0)[0])
    arr = np.asarray(chains, dtype=np.float64) if isinstance(chains, (list, tuple)) else chains
    mean, var, N = chain_moments(arr, device, draws_first)
    scalar = mean.dim() <= 1
    mean = mean.reshape(mean.shape[0] if mean.dim() else 1, -1)
    var = var.reshape(mean.shape)
    if D_.rank_world(group)[1] > 1:
        mean, var = D_.all_gather_cat(mean, group), D_.all_gather_cat(var, group)
    out = _rhat_from(mean, var, None, N)
    if scalar:
        return float(out[0]) if host else out[0]
    return out
