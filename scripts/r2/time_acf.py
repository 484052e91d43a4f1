"""autocorr timing: 8192 series x N draws (fp32 in, fp64 out): python scripts/r2/time_acf.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
for N in [int(v) for v in sys.argv[1:]] or [10000]:
    S = 8192
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn(S, N, device="cuda", generator=g).cumsum(1) * 0.01 + torch.randn(S, N, device="cuda", generator=g)
    bk.autocorr(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = bk.autocorr(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"N={N}: {ms:.3f} ms for {S} series = {S / ms / 1e3:.3f} M series/s  (check {float(out[:, 1].double().mean()):.6f})", flush=True)
