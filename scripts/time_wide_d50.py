"""A/B helper: fused samplers at D = 50 (BK_SEP_WIDE=1: 4 lanes x 16 elements per chain, 0: 16 x 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
C, D, n = 1048576, 50, 10
for name, s in (("HMCDiag L=10", bk.HMCDiag(bk.IsoGauss(D), 0.1, 10, chains=C, seed=0)),
                ("MALA", bk.MALA(bk.IsoGauss(D), 0.05, chains=C, seed=0)),
                ("DrGhmcDiag K=2", bk.DrGhmcDiag(bk.IsoGauss(D), 2, [0.6, 0.2], [5, 10], 0.5, chains=C // 4, seed=0))):
    ms = timed(lambda: s.sample_n(n))
    cc = C if "Dr" not in name else C // 4
    print(f"wide={os.environ.get('BK_SEP_WIDE', '1')} D={D} {name}: {cc * n / ms / 1e6:.2f} G chain-steps/s", flush=True)
