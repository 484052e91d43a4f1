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
def both(up, full):
    """ramp-up chunks, full chunks, mirrored ramp-down; the remainder joins the middle"""
    down = list(reversed(up))
    mid = C - sum(up) - sum(down)
    n = max(1, round(mid / full))
    base = (mid // n) // 256 * 256
    mids = [base] * n
    mids[n // 2] += mid - base * n
    return list(up) + mids + down
layouts = {"9472": 9472,
           "ramp-up 2304,4608 + 9472": ramp(2304, 4608, full=9472),
           "both 2304,4608 | 9472": both([2304, 4608], 9472),
           "both 1280,2304,4864 | 9472": both([1280, 2304, 4864], 9472),
           "both 1024,2048,4096 | 8192": both([1024, 2048, 4096], 8192),
           "both 768,1536,3072,6144 | 12288": both([768, 1536, 3072, 6144], 12288),
           "both 512,1024,2048,4096 | 8192": both([512, 1024, 2048, 4096], 8192)}
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

# timeline of one step with the default layout: start / end of every chunk's copy-in, kernels and copy-out (ms from
# the first event), from CUDA events recorded on the three streams
s._trace = []
s.sample_host(bufs[0], out=(bufs[1], lp))
torch.cuda.synchronize()
ref = min((t for t in s._trace if t[0] == "h2d"), key=lambda t: t[1])[2]
for kind in ("h2d", "run", "d2h"):
    print(kind, " ".join(f"[{ref.elapsed_time(a):.2f}-{ref.elapsed_time(b):.2f}]" for k, _, a, b in s._trace if k == kind))
s._trace = None
