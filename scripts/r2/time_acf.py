"""autocorr timing: 8192 series x N draws (fp32 in, fp64 out): python scripts/r2/time_acf.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
for N in [int(v) for v in sys.argv[1:]] or [10000]:
    S = 8192
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn(S, N, device="cuda", generator=g).cumsum(1) * 0.01 + torch.randn(S, N, device="cuda", generator=g)
    bk.autocorr(x); bk.autocorr(x); torch.cuda.synchronize()
    times = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = bk.autocorr(x); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    print(f"N={N}: median {ms:.3f} ms for {S} series = {S / ms / 1e3:.3f} M series/s  (check {float(out[:, 1].double().mean()):.6f}) "
          f"reps {' '.join(f'{t:.2f}' for t in times)}", flush=True)
