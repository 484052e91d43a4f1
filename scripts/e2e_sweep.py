"""sample_host() chunking sweep on the c2 workload: ms per step for several chunk layouts (host wall clock around
the call; sample_host returns with the draw complete in pinned host memory)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bayes_kit_b200 as bk
from oracle.models import DensePrecGauss
D, C, L = 1000, 65536, 10
model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
s = bk.HMCDiag(model, 0.1, L, chains=C, seed=0)
bufs = [torch.randn(C, D).pin_memory(), torch.empty(C, D).pin_memory()]
lp = torch.empty(C).pin_memory()
def ramp(*head, full):
    sizes = list(head)
    while sum(sizes) < C:
        sizes.append(min(full, C - sum(sizes)))
    return sizes
layouts = {"9472": 9472, "13056": 13056, "16384": 16384, "21760": 21760, "32768": 32768,
           "ramp 2304,4608 + 9472": ramp(2304, 4608, full=9472),
           "ramp 4096,8192 + 16384": ramp(4096, 8192, full=16384),
           "ramp 2048,4096,8192 + 16384": ramp(2048, 4096, 8192, full=16384),
           "ramp 4096,8192,16384 + 32768 (tail 4352)": [4096, 8192, 16384, 32512, 4352]}
for name, ch in layouts.items():
    for _ in range(2):
        s.sample_host(bufs[0], out=(bufs[1], lp), chunk_chains=ch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 6
    for i in range(n):
        s.sample_host(bufs[i % 2], out=(bufs[(i + 1) % 2], lp), chunk_chains=ch)
    dt = (time.perf_counter() - t0) / n
    print(f"{name:45s} {dt * 1e3:6.2f} ms/step  {C / dt / 1e6:6.2f} M chain-steps/s", flush=True)
