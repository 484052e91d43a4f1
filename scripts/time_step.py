"""STEP-launch timing of the c2 pipeline (in-library CUDA events); BK_TC_DEBUG / BK_TC_PAIR are honoured,
so `for d in 0 1 2; do BK_TC_DEBUG=$d python scripts/time_step.py; done` isolates GEMM and epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib
from oracle.models import DensePrecGauss
lib = _lib.lib()
D, C, L = 1000, int(os.environ.get("PROF_C", "65536")), 10
model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
s = bk.HMCDiag(model, 0.1, L, chains=C, seed=0)
s.sample_n(3)
torch.cuda.synchronize()
lib.bk_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.sample_n(10); s.sample_n(10); s.sample_n(10); e1.record(); torch.cuda.synchronize()
ms, nn = _lib.f64(0), _lib.u64(0)
lib.bk_profile_read(_lib.PROF_STEP, ms, nn)
print(f"debug={os.environ.get('BK_TC_DEBUG', '0')} pair={os.environ.get('BK_TC_PAIR', '1')}: "
      f"{e0.elapsed_time(e1) / 30:.3f} ms/draw, STEP {ms.value / 30 / (L - 1) * 1e3:.1f} us per leapfrog step ({nn.value} launches)",
      flush=True)
