"""Shared host logic of the lockstep MCMC samplers (one object == C chains)."""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from . import _lib as L
from ._util import Workspace, make_rng, ptr, resolve_seed, to_dev
from .models import require_plugin

DrawAndLogP = Tuple[torch.Tensor, torch.Tensor]


def _is_empty_init(init) -> bool:
    # reference: `init is not None and init.shape != (0,)` (hmc.py:24-28)
    return init is None or tuple(getattr(init, "shape", (1,))) == (0,)


class ChainSampler:
    """C independent chains advanced in lockstep on one GPU.

    ``init`` of shape [D] keeps the reference's single-chain surface
    (``sample() -> (theta[D], logp)``); shape [C, D] (or ``chains=C``) runs C
    chains and ``sample()`` returns ``(Theta[C, D], logp[C])``.  Randomness is
    device Philox keyed by (seed, global chain id, draw index) -- results do
    not depend on how chains are sharded over GPUs -- unless pre-drawn streams
    are injected (``sample_n(n, normals=..., uniforms=...)``, parity mode).
    """

    _n_uniform = 1

    def __init__(self, model, init, seed, chains: Optional[int], chain_offset: int = 0):
        self._model = require_plugin(model)
        self._dim = self._model.dims()
        self.device, self.dtype = self._model.device, self._model.dtype
        self._seed = resolve_seed(seed)
        self._chain_offset = int(chain_offset)
        self._t = 0  # draws taken so far (Philox draw counter)
        if _is_empty_init(init):
            n = 1 if chains is None else int(chains)
            self._single = chains is None
            # theta0 ~ N(0, I) like rng.normal(size=dim) (hmc.py:27)
            g = torch.Generator(device=self.device)
            g.manual_seed((self._seed * 0x9E3779B97F4A7C15 + self._chain_offset) % (2**63))
            th = torch.randn(n, self._dim, generator=g, device=self.device, dtype=self.dtype)
        else:
            th = to_dev(init, self.dtype, self.device)
            self._single = th.dim() == 1
            th = th.reshape(1, -1) if self._single else th
            if th.dim() != 2 or th.shape[1] != self._dim:
                raise ValueError(f"init must have shape [{self._dim}] or [C, {self._dim}], "
                                 f"got {tuple(th.shape)}")
            if chains is not None and int(chains) != th.shape[0]:
                raise ValueError("chains does not match init.shape[0]")
            th = th.clone()
        self._theta = th.contiguous()
        self._C = self._theta.shape[0]
        self._ws = Workspace(self.device)
        self._lp = torch.empty(self._C, dtype=self.dtype, device=self.device)
        self._grad = None
        self._cache_valid = L.i32(0)
        self.last_accept = None

    # ---- reference surface ---------------------------------------------------------
    def __iter__(self) -> Iterator[DrawAndLogP]:
        return self

    def __next__(self) -> DrawAndLogP:
        return self.sample()

    def sample(self) -> DrawAndLogP:
        draws, logp = self.sample_n(1)
        return draws[0], logp[0]

    # ---- batched extension -----------------------------------------------------------
    @property
    def chains(self) -> int:
        return self._C

    @property
    def theta(self) -> torch.Tensor:
        return self._theta[0] if self._single else self._theta

    def sample_n(self, n: int, normals=None, uniforms=None, keep_draws: bool = True):
        """Advance every chain n draws in one call.  Returns (draws [n, C, D],
        logp [n, C]) ([n, D], [n] for a single chain); with keep_draws=False only
        the final state is kept (warm-up) and draws is None."""
        n = int(n)
        C_, D = self._C, self._dim
        draws = torch.empty(n, C_, D, dtype=self.dtype, device=self.device) if keep_draws else None
        logp = torch.empty(n, C_, dtype=self.dtype, device=self.device)
        acc = torch.empty(n, C_, dtype=torch.int32, device=self.device)
        if normals is not None:
            normals = to_dev(normals, self.dtype, self.device).reshape(n, C_, D)
            uniforms = to_dev(uniforms, self.dtype, self.device).reshape(n, C_, self._n_uniform)
        rng = make_rng(self._seed, self._t, self._chain_offset, normals, uniforms, self._n_uniform)
        out = L.DrawOut(ptr(draws), logp.data_ptr(), acc.data_ptr())
        with torch.cuda.device(self.device):
            self._launch(n, rng, out)
        self._t += n
        self.last_accept = acc[:, 0] if self._single else acc
        if self._single:
            return (draws[:, 0] if keep_draws else None), logp[:, 0]
        return draws, logp

    def _launch(self, n, rng, out):  # pragma: no cover - overridden
        raise NotImplementedError

    def _need_grad_cache(self):
        if self._grad is None:
            self._grad = torch.empty_like(self._theta)
