"""Small-shape pass over the kernels changed in the second session of round 2 (for compute-sanitizer memcheck):
narrow-layout fused samplers, turn kernel (vector and scalar rows, L = 1 / 2 / 5, MALA), third position array of the
STEP launch, SMC move with compile-time row access modes, on-chip rfft autocorrelation (several plans), the
multi-step regression launch (BK_HLR_FUSE=1 in the environment selects it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bayes_kit_b200 as bk
from oracle.models import DensePrecGauss, HierLogReg

for D in (70, 84, 100, 107, 50, 33):
    for mk in (lambda: bk.HMCDiag(bk.IsoGauss(D), 0.15, 5, chains=333, seed=1),
               lambda: bk.HMCDiag(bk.IsoGauss(D), 0.15, 4, chains=333, seed=1),
               lambda: bk.MALA(bk.IsoGauss(D), 0.05, chains=333, seed=2),
               lambda: bk.Metropolis(bk.IsoGauss(D), bk.GaussianRW(0.2), chains=333, seed=3)):
        s = mk(); s.sample_n(3); s.sample_n(2, keep_draws=False, moments=True); s.sample_n(8, layout="series")
for C, D, L in ((700, 1000, 5), (300, 250, 2), (257, 1001, 1), (512, 128, 3)):
    model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
    s = bk.HMCDiag(model, 0.1, L, chains=C, seed=3); s.sample_n(4); s.sample_n(1); s.sample_n(2)
    m = bk.MALA(model, 2e-3, chains=C, seed=4); m.sample_n(3)
for D in (50, 52, 45, 64):
    M, T = 5000, 5
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = torch.randn(D, device="cuda", generator=g)
    th0 = torch.randn(M, D, device="cuda", generator=g)
    model = bk.GaussPriorLik(torch.zeros(D), torch.ones(D), mu, 4 * torch.ones(D))
    for mode in ("systematic", "multinomial"):
        smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample=mode, seed=7); smc.run()
for N in (10000, 5000, 1000, 777, 300, 4097):
    x = torch.randn(37, N, device="cuda")
    bk.autocorr(x); bk.autocorr(x.double())
    bk.autocorr(torch.randn(N, 5, 3, device="cuda"), draws_first=True)
X, y = HierLogReg.c3_data(3000, 37, seed=0)
s = bk.HMCDiag(bk.HierLogReg(X, y), 0.01, 4, init=np.random.default_rng(1).normal(size=(300, 39)) * 0.1, seed=0)
s.sample_n(2)
torch.cuda.synchronize()
print("sanitize_small done")
