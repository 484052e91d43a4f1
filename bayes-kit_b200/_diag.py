"""Host glue shared by the diagnostics: array-likes -> device series layouts."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._util import Workspace, require_cuda, stream_ptr

_WS = {}


def as_series(chain, device="cuda", draws_first=False):
    """-> (device tensor, result_shape, SeriesLayout).

    Default layouts: 1-D [N]; 2-D [chains, draws]; 3-D [chains, draws, params].
    ``draws_first=True``: [draws, chains] / [draws, chains, params] -- the
    layout the samplers write -- consumed in place (no transpose).
    result_shape is the shape of a per-series result ((), [chains] or
    [chains, params])."""
    if isinstance(chain, torch.Tensor) and chain.is_cuda:
        x = chain
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        x = x.contiguous()
    else:
        dev = require_cuda(device)
        arr = np.asarray(chain.cpu() if isinstance(chain, torch.Tensor) else chain, dtype=np.float64)
        x = torch.as_tensor(arr, device=dev)
    sh = tuple(x.shape)
    lay = L.SeriesLayout()
    if len(sh) == 1:
        res, vals = (), (1, sh[0], 1, sh[0], 0, 1)
    elif len(sh) == 2 and not draws_first:
        res, vals = (sh[0],), (sh[0], sh[1], 1, sh[1], 0, 1)
    elif len(sh) == 3 and not draws_first:
        c, n, p = sh
        res, vals = (c, p), (c * p, n, p, n * p, 1, p)
    elif len(sh) == 2:
        n, c = sh
        res, vals = (c,), (c, n, c, 0, 1, c)
    elif len(sh) == 3:
        n, c, p = sh
        res, vals = (c, p), (c * p, n, c * p, 0, 1, c * p)
    else:
        raise ValueError("expected a 1-D, 2-D or 3-D array of draws")
    (lay.n_series, lay.n_draws, lay.n_inner, lay.outer_stride, lay.inner_stride,
     lay.draw_stride) = vals
    return x, res, lay


def call(x, fn, ws_bytes=0):
    """Run `fn(lib, x_ptr, dtype_id, stream, ws_ptr, ws_bytes)` on x's device."""
    lib = L.lib()
    dev = x.device
    ws = _WS.setdefault(dev, Workspace(dev))
    wp, wn = ws.get(ws_bytes)
    dt = L.BK_F32 if x.dtype == torch.float32 else L.BK_F64
    with torch.cuda.device(dev):
        L.check(fn(lib, x.data_ptr(), dt, stream_ptr(dev), wp, wn))


def finish(out, res_shape, host_input):
    """Per-series results -> reference-shaped return (float for a 1-D host chain)."""
    if res_shape == ():
        return float(out[0]) if host_input else out[0]
    return out.reshape(res_shape)


def is_host(chain) -> bool:
    return not (isinstance(chain, torch.Tensor) and chain.is_cuda)
