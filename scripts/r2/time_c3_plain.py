"""c3 draw time without the in-library event profiling (which serialises programmatic dependent launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayes_kit_b200 as bk
from oracle.models import HierLogReg
N, Dx, C = 100_000, 100, 1024
X, y = HierLogReg.c3_data(N, Dx, seed=0)
s = bk.HMCDiag(bk.HierLogReg(X, y), 0.01, 10, init=np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1, seed=0)
s.sample_n(3); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d, lp = s.sample_n(10); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
import hashlib
print(f"PDL={os.environ.get('BK_HLR_PDL', '1')}: median {sorted(ts)[2]:.4f} ms/draw  accept {float(s.last_accept.float().mean()):.3f}  "
      f"hash {hashlib.sha1(d.cpu().numpy().tobytes()).hexdigest()[:12]}", flush=True)
