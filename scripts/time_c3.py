"""c3 timings: HMCDiag on the hierarchical logistic regression, per-L breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib
from oracle.models import HierLogReg
lib = _lib.lib()
N, Dx, C = 100_000, 100, int(os.environ.get("PROF_C", "1024"))
X, y = HierLogReg.c3_data(N, Dx, seed=0)
model = bk.HierLogReg(X, y)
th0 = np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1
for L in (1, 2, 10):
    s = bk.HMCDiag(model, 0.01, L, init=th0, seed=0)
    s.sample_n(2); torch.cuda.synchronize()
    lib.bk_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample_n(5); e1.record(); torch.cuda.synchronize()
    ms, nn = _lib.f64(0), _lib.u64(0)
    lib.bk_profile_read(_lib.PROF_GRAD, ms, nn); lib.bk_profile_enable(0)
    dt = e0.elapsed_time(e1) / 5
    print(f"L={L}: {dt:.3f} ms/draw; grad kernels/draw {nn.value/5:.0f} total {ms.value/5:.3f} ms; "
          f"accept {float(s.last_accept.float().mean()):.3f}; {C/dt*1e3:.0f} chain-steps/s", flush=True)
