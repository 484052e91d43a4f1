"""TemperedLikelihoodSMC (reference: bayes_kit/smc.py).

All M particles move, reweight and resample on device, one temperature per
``transition(n)``: a fused RW-Metropolis move + log-weight kernel, a block-scan
CDF + binary-search resampler and a row gather.  Particles shard across ranks;
the only communication is the naturally global resampling step (all-gather of
log-weights and particles over NCCL), see dist.py.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional

import numpy as np
import torch

from . import _lib as L
from . import dist as D_
from ._util import Workspace, make_rng, ptr, resolve_seed, stream_ptr, to_dev
from .models import GaussPriorLik, require_plugin


class RWMetropolisKernel:
    """Device descriptor returned by ``metropolis_kernel(scale)``: one
    random-walk Metropolis move ``theta* = normal(loc=theta, scale)`` accepted
    iff ``log(u) < lp(theta*) - lp(theta)`` (smc.py:79-89)."""

    def __init__(self, scale: float):
        self.scale = float(scale)
        if not self.scale > 0:
            raise ValueError(f"scale must be positive, got {scale}")


def metropolis_kernel(scale: float) -> RWMetropolisKernel:
    return RWMetropolisKernel(scale)


class TemperedLikelihoodSMC:
    """``TemperedLikelihoodSMC(model, M, N, sample_initial, kernel)`` (smc.py:13-27).

    * ``model``: a ``GaussPriorLik`` plugin (``log_prior`` / ``log_likelihood``).
    * ``sample_initial``: the reference's callable ``m -> theta0[m]`` (evaluated
      once per particle on the host, as smc.py:23 does), or directly an [M, D]
      array / tensor.
    * ``kernel``: ``metropolis_kernel(scale)``.
    * ``resample``: ``"multinomial"`` (reference semantics: no max-shift, legacy
      ``np.random.choice`` indices) or ``"systematic"`` (log-sum-exp normalised,
      stratified points; north_star item 2).
    * ``ess_threshold``: None (the reference: resample at every temperature,
      smc.py:60) or a fraction in (0, 1]: resample only when the importance-weight
      ESS drops below ``ess_threshold * M``, carrying the log-weights otherwise
      (``log_weights``, ``resampled``); the decision is taken on device.
    With torch.distributed initialised, M is the GLOBAL particle count and each
    rank owns a contiguous slice.
    """

    def __init__(self, model, M: int, N: int, sample_initial, kernel, *, resample: str = "multinomial",
                 ess_threshold: Optional[float] = None, seed=None, group=None):
        self._model = require_plugin(model)
        if not isinstance(self._model, GaussPriorLik):
            raise TypeError("TemperedLikelihoodSMC needs a model plugin with log_prior/log_likelihood "
                            "(bayes_kit_b200.models.GaussPriorLik)")
        if not isinstance(kernel, RWMetropolisKernel):
            raise TypeError("kernel must be bayes_kit_b200.metropolis_kernel(scale): arbitrary Python "
                            "kernels cannot run per particle on the GPU and there is no CPU fallback")
        if resample not in ("multinomial", "systematic"):
            raise ValueError("resample must be 'multinomial' or 'systematic'")
        if ess_threshold is not None and not 0 < ess_threshold <= 1:
            raise ValueError(f"ess_threshold must be in (0, 1], got {ess_threshold}")
        self.ess_threshold = ess_threshold      # None: resample at every temperature (smc.py:60)
        self.M, self.N = int(M), int(N)
        self.kernel = kernel
        self.resample = resample
        self.device, self.dtype = self._model.device, self._model.dtype
        self._seed = resolve_seed(seed)
        self._group = group
        self._rank, self._world = D_.rank_world(group)
        self._lo, self._hi = D_.shard_range(self.M, self._rank, self._world)
        if callable(sample_initial):
            th = np.array([np.asarray(sample_initial(m), dtype=np.float64)
                           for m in range(self._lo, self._hi)])
            th = th.reshape(self._hi - self._lo, -1)
        else:
            th = sample_initial
            if th.shape[0] == self.M and self._world > 1:
                th = th[self._lo:self._hi]
        self._thetas = to_dev(th, self.dtype, self.device).clone()
        self._pending = None   # (all-gathered particles, resample indices) not yet materialised
        self.D = self._thetas.shape[1]
        if self.D != self._model.dims():
            raise ValueError("sample_initial returned the wrong dimension")
        self._ws = Workspace(self.device)
        self.last_indices = None
        self.last_accept = None
        self._stats_log = []  # per temperature: device tensor [shift, sum w, sum w^2]
        self._logw_acc = None  # adaptive resampling: log-weights carried between temperatures
        self._resampled_log = []

    # ---- reference surface -----------------------------------------------------------
    @property
    def thetas(self) -> torch.Tensor:
        """Current particles [M_local, D] (smc.py:23,60).  The resampling gather
        ``thetas[idxs]`` (smc.py:75) is normally folded into the next move's read;
        it is materialised here when the particles are asked for."""
        if self._pending is not None:
            src, idx = self._pending
            new = torch.empty(idx.shape[0], self.D, dtype=self.dtype, device=self.device)
            with torch.cuda.device(self.device):
                L.check(L.lib().bk_gather_rows(src.data_ptr(), idx.data_ptr(), idx.shape[0], self.D,
                                               L.BK_F32 if self.dtype == torch.float32 else L.BK_F64,
                                               new.data_ptr(), stream_ptr(self.device)))
            self._thetas, self._pending = new, None
        return self._thetas

    @thetas.setter
    def thetas(self, value) -> None:
        self._thetas, self._pending = to_dev(value, self.dtype, self.device).clone(), None

    def log_prior(self, theta):
        return self._model.log_prior(theta)

    def log_likelihood(self, theta):
        return self._model.log_likelihood(theta)

    def __iter__(self) -> Iterator[torch.Tensor]:
        self.run()
        return iter(self.thetas)

    def run(self) -> None:
        for n in range(1, self.N + 1):
            self.transition(n)

    def time(self, n: int) -> float:
        return n / self.N

    @property
    def weight_ess(self):
        """Importance-weight ESS (sum w)^2 / sum w^2 per temperature taken so far.
        Diagnostic only: the reference resamples unconditionally (smc.py:60)."""
        if not self._stats_log:
            return []
        s = torch.stack(self._stats_log).cpu()
        return [float(a * a / b) if b > 0 else float("nan") for _, a, b in s.tolist()]

    @property
    def log_weights(self) -> torch.Tensor:
        """Unnormalised log-weights of the current particles (zeros right after a resampling)."""
        if self._logw_acc is None:
            return torch.zeros(self._hi - self._lo, dtype=self.dtype, device=self.device)
        return self._logw_acc

    @property
    def resampled(self):
        """Per temperature taken so far: did the adaptive rule resample?"""
        if not self._resampled_log:
            return []
        return [bool(v) for v in torch.cat(self._resampled_log).cpu().tolist()]

    def transition(self, n: int, normals=None, acc_uniforms=None, res_uniforms=None) -> None:
        """One temperature step (smc.py:46-60).  The optional arrays inject the
        reference's recorded legacy-RNG streams (parity mode): proposal normals
        [M, D], accept uniforms [M], resampling uniforms [M] (systematic: [1])."""
        lib = L.lib()
        Ml = self._hi - self._lo
        st = stream_ptr(self.device)
        logw = torch.empty(Ml, dtype=self.dtype, device=self.device)
        acc = torch.empty(Ml, dtype=torch.int32, device=self.device)
        if normals is not None:
            normals = to_dev(normals, self.dtype, self.device).reshape(Ml, self.D)
            acc_uniforms = to_dev(acc_uniforms, self.dtype, self.device).reshape(Ml)
        rng = make_rng(self._seed, n, self._lo, normals, acc_uniforms, 1)
        mode = L.RESAMPLE_MULTINOMIAL if self.resample == "multinomial" else L.RESAMPLE_SYSTEMATIC
        with torch.cuda.device(self.device):
            prev = None if self._logw_acc is None else self._logw_acc.data_ptr()
            if self._pending is not None:     # move reads thetas_prev[idx]: gather + move in one pass
                src, src_idx = self._pending
                moved = torch.empty(Ml, self.D, dtype=self.dtype, device=self.device)
                L.check(lib.bk_smc_gather_move_weight_acc(
                    self._model.handle, src.data_ptr(), src_idx.data_ptr(), moved.data_ptr(), Ml, n, self.N,
                    self.kernel.scale, C.byref(rng), prev, logw.data_ptr(), acc.data_ptr(), st))
                self._thetas, self._pending = moved, None
            else:
                L.check(lib.bk_smc_gather_move_weight_acc(
                    self._model.handle, self._thetas.data_ptr(), None, self._thetas.data_ptr(), Ml, n, self.N,
                    self.kernel.scale, C.byref(rng), prev, logw.data_ptr(), acc.data_ptr(), st))
            # --- the naturally global step: normaliser + resampling ---------------------
            logw_all = D_.all_gather_cat(logw, self._group)      # [M]
            thetas_all = D_.all_gather_cat(self._thetas, self._group)  # [M, D]
            Mg = logw_all.shape[0]
            wp, wn = self._ws.get(lib.bk_smc_resample_workspace_bytes(Mg))
            stats = torch.empty(3, dtype=torch.float64, device=self.device)
            L.check(lib.bk_smc_weight_stats(logw_all.data_ptr(), Mg, L.BK_F32 if self.dtype == torch.float32
                                            else L.BK_F64, mode, stats.data_ptr(), wp, wn, st))
            self._stats_log.append(stats)   # stays on device: no host sync per temperature
            if res_uniforms is not None:
                ru = to_dev(res_uniforms, self.dtype, self.device).reshape(-1)
                if mode == L.RESAMPLE_MULTINOMIAL and ru.numel() == Mg and self._world > 1:
                    ru = ru[self._lo:self._hi].contiguous()
            else:
                ru = None
            idx = torch.empty(Ml, dtype=torch.int64, device=self.device)
            rrng = make_rng(self._seed, n, 0)
            L.check(lib.bk_smc_resample_indices_dev(
                logw_all.data_ptr(), Mg, L.BK_F32 if self.dtype == torch.float32 else L.BK_F64, mode,
                stats.data_ptr(), ptr(ru), C.byref(rrng), Ml, self._lo, idx.data_ptr(), None, wp, wn, st))
            if self.ess_threshold is not None:
                flag = torch.empty(1, dtype=torch.int32, device=self.device)
                L.check(lib.bk_smc_adaptive_select(
                    stats.data_ptr(), float(self.ess_threshold) * Mg, Ml, self._lo, idx.data_ptr(), logw.data_ptr(),
                    L.BK_F32 if self.dtype == torch.float32 else L.BK_F64, flag.data_ptr(), st))
                self._logw_acc = logw
                self._resampled_log.append(flag)
        self._pending = (thetas_all, idx)     # thetas[idxs]: folded into the next move (or .thetas)
        self.last_indices = idx
        self.last_accept = acc
