"""ess timing: 25,600 AR(1) series x 10,000 draws, series-major fp32 (the c5 leg's shape): python scripts/r2/time_ess.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
N, S = 10000, 25600
g = torch.Generator(device="cuda").manual_seed(0)
phi = torch.rand(S, device="cuda", generator=g) * 0.9
x = torch.empty(N, S, device="cuda"); cur = torch.randn(S, device="cuda", generator=g)
for t in range(N):
    cur = phi * cur + torch.randn(S, device="cuda", generator=g); x[t] = cur
xs = x.t().contiguous()
for name, arr in (("series-major", xs), ("series-major, misaligned view", xs[:, 1:])):
    bk.ess(arr); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e = bk.ess(arr); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[3]
    print(f"{name}: {ms:.3f} ms = {S / ms / 1e3:.2f} M series/s  (mean ess {float(e.mean()):.3f})", flush=True)
