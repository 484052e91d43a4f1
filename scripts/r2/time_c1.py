"""A/B helper: fused samplers at D = 100 under the lane layouts of the isotropic fp32 plugin
(BK_SEP_LAYOUT=0: 8 lanes x 4 blocks, 4: 4 x 7, 2: 2 x 13).  Prints rate, accept rate, pooled moments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk


def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


C, n = 1048576, 10
D = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lay = os.environ.get("BK_SEP_LAYOUT", "default")
for name, mk in (("HMCDiag L=10", lambda: bk.HMCDiag(bk.IsoGauss(D), 0.1, 10, chains=C, seed=0)),
                 ("HMCDiag L=9", lambda: bk.HMCDiag(bk.IsoGauss(D), 0.1, 9, chains=C, seed=0)),
                 ("MALA", lambda: bk.MALA(bk.IsoGauss(D), 0.05, chains=C, seed=0)),
                 ("Metropolis", lambda: bk.Metropolis(bk.IsoGauss(D), bk.GaussianRW(0.2), chains=C, seed=0))):
    s = mk()
    ms = timed(lambda: s.sample_n(n))
    d, lp = s.sample_n(4)
    x = d[-1].double()
    print(f"layout={lay} D={D} {name}: {C * n / ms / 1e6:.3f} G chain-steps/s  accept={float(s.last_accept.float().mean()):.4f} "
          f"mean={float(x.mean()):+.5f} var={float(x.var()):.5f} sum={float(d.double().sum()):.6f}", flush=True)
