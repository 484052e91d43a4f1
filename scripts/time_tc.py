"""Per-kernel timings of the c2 pipeline (device-resident, CUDA events in-library)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib
from oracle.models import DensePrecGauss
lib = _lib.lib()
D, C = 1000, int(os.environ.get("PROF_C", "65536"))
model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
for L in (1, 2, 10):
    s = bk.HMCDiag(model, 0.1, L, chains=C, seed=0)
    s.sample_n(3)
    torch.cuda.synchronize()
    lib.bk_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample_n(10); e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 10
    ms, nn = _lib.f64(0), _lib.u64(0)
    lib.bk_profile_read(_lib.PROF_GRAD, ms, nn)
    lib.bk_profile_enable(0)
    print(f"L={L}: {dt:.3f} ms/draw; gemm launches/draw {nn.value/10:.0f}, {ms.value/10:.3f} ms/draw, "
          f"avg {ms.value/max(nn.value,1):.3f} ms; rest {dt-ms.value/10:.3f} ms; "
          f"{C/dt/1e3:.2f} M chain-steps/s", flush=True)
