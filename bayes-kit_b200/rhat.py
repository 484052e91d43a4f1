"""Potential scale reduction (reference: bayes_kit/rhat.py:111-171)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import dist as D_
from ._diag import as_series, call, is_host
from ._util import Workspace, require_cuda, stream_ptr


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


def _rhat_allreduce(mean, var, N, group=None, reduce_fn=None):
    """R-hat over chains sharded across ranks: two all-reduces of [P, 2] columns of a [P, 4] sum table
    (count + sum of chain means -> grand mean; squared deviations + sum of chain variances), i.e.
    4 doubles per parameter on the wire in total -- never the per-chain moments (fewer than two chains
    in total give NaN: a rank cannot know the global count before the reduce).  ``reduce_fn(t)`` replaces
    ``dist.all_reduce`` (single-GPU "fake world" tests sum the ranks' tables on the host side)."""
    import torch.distributed as dist
    if reduce_fn is None:
        def reduce_fn(t):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    lib = L.lib()
    n_chains, n_params = mean.shape
    mean, var = mean.contiguous(), var.contiguous()
    f64 = dict(dtype=torch.float64, device=mean.device)
    sums, sqdev = torch.zeros(n_params, 3, **f64), torch.zeros(n_params, **f64)
    ref, out = torch.empty(n_params, **f64), torch.empty(n_params, **f64)
    with torch.cuda.device(mean.device):
        st = stream_ptr(mean.device)
        L.check(lib.bk_rhat_partial_sums(mean.data_ptr(), var.data_ptr(), n_chains, n_params, None, sums.data_ptr(), st))
        reduce_fn(sums)
        L.check(lib.bk_rhat_from_sums(sums.data_ptr(), None, n_params, int(N), ref.data_ptr(), None, st))
        L.check(lib.bk_rhat_partial_sums(mean.data_ptr(), None, n_chains, n_params, ref.data_ptr(), sqdev.data_ptr(), st))
        reduce_fn(sqdev)
        L.check(lib.bk_rhat_from_sums(sums.data_ptr(), sqdev.data_ptr(), n_params, int(N), None, out.data_ptr(), st))
    return out


def rhat(chains, device="cuda", draws_first=False, group=None, reduce_fn=None):
    """R-hat = sqrt((nbar-1)/nbar + var(chain means, ddof=1) / mean(chain vars, ddof=1)).

    ``chains``: the reference's list of 1-D chains (ragged allowed, returns a
    float), a ``[chains, draws]`` array (float / 0-dim tensor) or
    ``[chains, draws, params]`` (tensor [params]).  With ``group`` given
    (``torch.distributed.group.WORLD`` for all ranks; the call is then COLLECTIVE over that
    group -- without ``group`` it is always local, also under torchrun), ``chains`` is this
    rank's shard of chains (all of one length) and the only communication is an
    all-reduce of a [params, 4] table of moment sums (``_rhat_allreduce``).  Raises ValueError for < 2 chains or a chain with < 2
    draws (rhat.py:157-162)."""
    host = is_host(chains)
    ragged = (isinstance(chains, (list, tuple)) and len(chains) > 0
              and len({len(c) for c in chains}) > 1)
    if isinstance(chains, (list, tuple)):
        if len(chains) < 2 and (group is None or D_.rank_world(group)[1] == 1):
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
    if (group is not None and D_.rank_world(group)[1] > 1) or reduce_fn is not None:
        out = _rhat_allreduce(mean, var, N, group, reduce_fn)
    else:
        out = _rhat_from(mean, var, None, N)
    if scalar:
        return float(out[0]) if host else out[0]
    return out


# ---- split / rank-normalised R-hat (reference: rhat.py:9-108, 174-236) -------------------
def _device_chains(chains, device, draws_first):
    """-> (x [chains, draws(, params)] device tensor in place (no copy for device input),
    host_input).  Lists of equal-length 1-D chains become a [chains, draws] array."""
    host = is_host(chains)
    if isinstance(chains, (list, tuple)):
        chains = np.asarray([np.asarray(c, dtype=np.float64) for c in chains], dtype=np.float64)
    x, _, _ = as_series(chains, device, draws_first=False) if not draws_first else as_series(chains, device, True)
    return x, host


def _layout_for(x, draws_first, p, d0, nd):
    """SeriesLayout of parameter p's chains restricted to draws [d0, d0 + nd), plus the
    element offset of its first value (x is [C, N(, P)] or, draws_first, [N, C(, P)])."""
    sh = tuple(x.shape)
    P = sh[2] if len(sh) == 3 else 1
    C_, N = (sh[1], sh[0]) if draws_first else (sh[0], sh[1])
    lay = L.SeriesLayout()
    if draws_first:      # element (t, c, p) at t*C*P + c*P + p
        vals, off = (C_, nd, 1, P, 0, C_ * P), d0 * C_ * P + p
    else:                # element (c, t, p) at c*N*P + t*P + p
        vals, off = (C_, nd, 1, N * P, 0, P), d0 * P + p
    (lay.n_series, lay.n_draws, lay.n_inner, lay.outer_stride, lay.inner_stride, lay.draw_stride) = vals
    return lay, off


def _moments_of(x, lay, off):
    m = torch.empty(lay.n_series, dtype=torch.float64, device=x.device)
    v = torch.empty_like(m)
    dt = L.BK_F32 if x.dtype == torch.float32 else L.BK_F64
    with torch.cuda.device(x.device):
        L.check(L.lib().bk_chain_moments(x.data_ptr() + off * x.element_size(), dt, C.byref(lay), m.data_ptr(),
                                         v.data_ptr(), stream_ptr(x.device)))
    return m, v


def split_chains(chains):
    """rhat.py:9-24: every chain split in half, the first half one longer for odd
    sizes.  Host helper (views, no device work) -- the device path below never
    materialises the halves."""
    out = []
    for c in chains:
        h = (len(c) + 1) // 2
        out.extend([c[:h], c[h:]])
    return out


def split_rhat(chains, device="cuda", draws_first=False):
    """``rhat(split_chains(chains))`` (rhat.py:174-202): the moments kernel runs on
    the two halves of every chain IN PLACE (pointer offset + draw count), then the
    ragged R-hat kernel.  ``chains``: list of 1-D chains (ragged allowed),
    ``[chains, draws]`` or ``[chains, draws, params]`` (-> tensor [params])."""
    if isinstance(chains, (list, tuple)):
        if len(chains) == 0:
            raise ValueError("rhat requires len(chains) >= 2, but len(chains) = 0")
        if len({len(c) for c in chains}) > 1:        # ragged: the plain ragged path on the halves
            return rhat(split_chains([np.asarray(c, dtype=np.float64) for c in chains]), device)
    x, host = _device_chains(chains, device, draws_first)
    if x.dim() == 1:
        x = x.reshape(1, -1)
    sh = tuple(x.shape)
    P = sh[2] if len(sh) == 3 else 1
    N = sh[0] if draws_first else sh[1]
    h1 = (N + 1) // 2
    if N - h1 < 2:
        raise ValueError("rhat requires len(chain) >= 2 for every chain in chains")
    C_ = sh[1] if draws_first else sh[0]

    def half(d0, nd):      # every (chain, param) series at once, restricted to draws [d0, d0 + nd)
        lay = L.SeriesLayout()
        if draws_first:    # element (t, c, p) at t*C*P + c*P + p
            vals, off = (C_ * P, nd, C_ * P, 0, 1, C_ * P), d0 * C_ * P
        else:              # element (c, t, p) at c*N*P + t*P + p
            vals, off = (C_ * P, nd, P, N * P, 1, P), d0 * P
        (lay.n_series, lay.n_draws, lay.n_inner, lay.outer_stride, lay.inner_stride, lay.draw_stride) = vals
        m, v = _moments_of(x, lay, off)
        return m.reshape(C_, P), v.reshape(C_, P)

    ma, va = half(0, h1)
    mb, vb = half(h1, N - h1)
    mean = torch.stack([ma, mb], 1).reshape(2 * C_, P)           # reference order: chain 0 halves, chain 1 halves, ...
    var = torch.stack([va, vb], 1).reshape(2 * C_, P)
    lengths = torch.tensor([h1, N - h1], dtype=torch.int64, device=x.device).repeat(C_)
    out = _rhat_from(mean, var, lengths, 0)
    if len(sh) == 3:
        return out
    return float(out[0]) if host else out[0]


def _rank_normalize_device(x, draws_first, want_ranks):
    """x [C, N(, P)] / [N, C(, P)] device tensor -> (ranks, z) float64 tensors [C, N(, P)]
    (chain-major, the reference's concatenation order)."""
    lib = L.lib()
    sh = tuple(x.shape)
    P = sh[2] if len(sh) == 3 else 1
    C_, N = (sh[1], sh[0]) if draws_first else (sh[0], sh[1])
    dt = L.BK_F32 if x.dtype == torch.float32 else L.BK_F64
    ws = Workspace(x.device)
    wp, wn = ws.get(lib.bk_rank_normalize_workspace_bytes(C_ * N, dt))
    z = torch.empty(P, C_, N, dtype=torch.float64, device=x.device)
    rk = torch.empty(P, C_, N, dtype=torch.float64, device=x.device) if want_ranks else None
    with torch.cuda.device(x.device):
        for p in range(P):
            lay, off = _layout_for(x, draws_first, p, 0, N)
            L.check(lib.bk_rank_normalize(x.data_ptr() + off * x.element_size(), dt, C.byref(lay),
                                          None if rk is None else rk[p].data_ptr(), z[p].data_ptr(), wp, wn,
                                          stream_ptr(x.device)))
    if len(sh) == 3:
        return (None if rk is None else rk.permute(1, 2, 0)), z.permute(1, 2, 0)
    return (None if rk is None else rk[0]), z[0]


def _ranked(chains, device, draws_first, want_ranks):
    ragged = isinstance(chains, (list, tuple)) and len({len(c) for c in chains}) > 1
    if ragged:     # ranks over the concatenation (rhat.py:51): one long series, cut back afterwards
        flat = np.concatenate([np.asarray(c, dtype=np.float64) for c in chains])
        x, _ = _device_chains(flat.reshape(1, -1), device, False)
        rk, z = _rank_normalize_device(x, False, want_ranks)
        cuts = np.cumsum([len(c) for c in chains])[:-1].tolist()
        split = lambda t: [a.cpu().numpy() for a in torch.tensor_split(t.reshape(-1), cuts)]
        return (None if rk is None else split(rk)), split(z), True
    x, host = _device_chains(chains, device, draws_first)
    if x.dim() == 1:
        x = x.reshape(1, -1)
    rk, z = _rank_normalize_device(x, draws_first, want_ranks)
    return rk, z, host


def rank_chains(chains, device="cuda", draws_first=False):
    """rhat.py:27-59: ranks (ascending from 1, float64) over the concatenation of the
    chains, in the shape of the input.  Lists come back as lists of NumPy arrays,
    device tensors as a device tensor ``[chains, draws(, params)]``.  Ties are ranked in
    flattened order (stable sort; the reference's tie order is implementation defined)."""
    if isinstance(chains, (list, tuple)) and len(chains) == 0:
        return chains
    rk, _, host = _ranked(chains, device, draws_first, True)
    if isinstance(rk, list):
        return rk
    return [r.cpu().numpy() for r in rk] if host else rk


def rank_normalize_chains(chains, device="cuda", draws_first=False):
    """rhat.py:62-108: ``norm.ppf((rank - 0.325) / (S - 0.25))`` for every draw."""
    if isinstance(chains, (list, tuple)) and len(chains) == 0:
        return []
    _, z, host = _ranked(chains, device, draws_first, False)
    if isinstance(z, list):
        return z
    return [r.cpu().numpy() for r in z] if host else z


def rank_normalized_rhat(chains, device="cuda", draws_first=False):
    """rhat.py:205-236: split R-hat of the rank-normalised chains; everything stays on
    device (radix sort -> normal scores -> half-chain moments -> R-hat)."""
    if isinstance(chains, (list, tuple)):
        if len(chains) == 0:
            raise ValueError("rhat requires len(chains) >= 2, but len(chains) = 0")
        if len({len(c) for c in chains}) > 1:
            return split_rhat(rank_normalize_chains(chains, device), device)
    host = is_host(chains)
    _, z, _ = _ranked(chains, device, draws_first, False)      # [C, N(, P)] float64, chain-major
    out = split_rhat(z, device)
    return float(out) if (host and not isinstance(out, float) and out.dim() == 0) else out
