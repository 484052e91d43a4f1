"""Metropolis / MetropolisHastings (reference: bayes_kit/metropolis.py).

The reference takes arbitrary Python callables ``proposal_fn(theta)`` and
``transition_lp_fn(to, from)`` (metropolis.py:79-88).  Python callbacks cannot
run per step on the device and there is no CPU fallback, so the proposal is a
built-in *device proposal descriptor*: ``GaussianRW(scale)`` -- the family the
reference's tests and SMC kernel use (test_metropolis.py:114, smc.py:79-89).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

from . import _lib as L
from ._sampler import ChainSampler
from ._util import stream_ptr


def metropolis_accept_test(lp_proposal: float, lp_current: float, rng) -> bool:
    """metropolis.py:12-38 (host scalar form, strict ``<``; ``log(0) = -inf``
    always accepts)."""
    u = rng.uniform()
    log_u = math.log(u) if u > 0 else -math.inf
    return log_u < lp_proposal - lp_current


def metropolis_hastings_accept_test(lp_proposal: float, lp_current: float,
                                    lp_forward_transition: float, lp_reverse_transition: float,
                                    rng) -> bool:
    """metropolis.py:41-76."""
    u = rng.uniform()
    log_u = math.log(u) if u > 0 else -math.inf
    return log_u < (lp_proposal - lp_current) + (lp_reverse_transition - lp_forward_transition)


class GaussianRW:
    """Symmetric Gaussian random-walk proposal ``theta' = theta + scale * z``."""

    def __init__(self, scale: float):
        self.scale = float(scale)
        if not self.scale > 0:
            raise ValueError(f"scale must be positive, got {scale}")

    def transition_lp(self, to, frm):  # marker, evaluated on device
        raise TypeError("GaussianRW.transition_lp is a device-side descriptor; pass it to "
                        "MetropolisHastings, do not call it")


def _require_rw(proposal_fn) -> GaussianRW:
    if not isinstance(proposal_fn, GaussianRW):
        raise TypeError(
            "proposal_fn must be a device proposal descriptor (bayes_kit_b200.GaussianRW(scale)): "
            "arbitrary Python callables cannot run per step on the GPU and there is no CPU fallback")
    return proposal_fn


class MetropolisHastings(ChainSampler):
    """``MetropolisHastings(model, proposal_fn, transition_lp_fn, *, init=None,
    seed=None)`` (metropolis.py:80-99).  ``transition_lp_fn`` must be the
    proposal's own ``transition_lp`` (the Hastings terms are evaluated on
    device; they cancel exactly for a symmetric proposal)."""

    _hastings = 1

    def __init__(self, model, proposal_fn, transition_lp_fn, *, init=None, seed=None,
                 chains: Optional[int] = None, chain_offset: int = 0):
        prop = _require_rw(proposal_fn)
        if self._hastings and getattr(transition_lp_fn, "__self__", None) is not prop:
            raise TypeError("transition_lp_fn must be proposal_fn.transition_lp (device descriptor)")
        super().__init__(model, init, seed, chains, chain_offset)
        self._proposal_fn = prop
        self._transition_lp_fn = transition_lp_fn

    def _launch(self, n, rng, out, c0=0, cn=None, cache_valid=None):
        lib = L.lib()
        cn = self._C if cn is None else cn
        wp, wn = self._ws.get(lib.bk_mh_rw_workspace_bytes(self._model.handle, self._C))
        L.check(lib.bk_mh_rw_sample(
            self._model.handle, self._theta[c0:].data_ptr(), self._lp[c0:].data_ptr(),
            C.byref(self._cache_valid if cache_valid is None else cache_valid), cn, self._proposal_fn.scale, self._hastings, n,
            C.byref(rng), C.byref(out), wp, wn, stream_ptr(self.device)))


class Metropolis(MetropolisHastings):
    """``Metropolis(model, proposal_fn, *, init=None, seed=None)``
    (metropolis.py:138-155): symmetric proposal, no Hastings term."""

    _hastings = 0

    def __init__(self, model, proposal_fn, *, init=None, seed=None, chains: Optional[int] = None,
                 chain_offset: int = 0):
        super().__init__(model, proposal_fn, None, init=init, seed=seed, chains=chains,
                         chain_offset=chain_offset)
