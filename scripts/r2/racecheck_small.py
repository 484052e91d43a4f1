"""Shared-memory kernels changed in the second session of round 2, small shapes (for compute-sanitizer racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
for N in (10000, 1000, 300):
    x = torch.randn(6, N, device="cuda")
    bk.autocorr(x); bk.ess(x)
bk.ess(torch.randn(500, 4, 3, device="cuda"), draws_first=True)
s = bk.HMCDiag(bk.IsoGauss(100), 0.15, 5, chains=200, seed=1); s.sample_n(9, layout="series")
torch.cuda.synchronize()
print("racecheck_small done")
