"""Per-launch time of the tcgen05 HLR gradient kernel vs number of observations (fixed cost vs per-tile cost)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib
from oracle.models import HierLogReg
lib = _lib.lib()
Dx = 100
for C in (1024, 4096):
    for N in (25_000, 50_000, 100_000, 200_000):
        X, y = HierLogReg.c3_data(N, Dx, seed=0)
        model = bk.HierLogReg(X, y)
        th = torch.as_tensor(np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1, dtype=torch.float32, device="cuda")
        for _ in range(3): model.log_density_gradient(th, fast=True)
        torch.cuda.synchronize()
        lib.bk_profile_enable(1)
        for _ in range(10): model.log_density_gradient(th, fast=True)
        torch.cuda.synchronize()
        ms, nn = _lib.f64(0), _lib.u64(0)
        lib.bk_profile_read(_lib.PROF_GRAD, ms, nn); lib.bk_profile_enable(0)
        t = ms.value / max(nn.value, 1)
        print(f"C={C} N={N}: {t*1e3:.1f} us/launch ({nn.value} launches), {4.0*N*Dx*C/t/1e9:.0f} TFLOP/s", flush=True)
