"""Small invocations of the fp32 kernels that use the 16-elements-per-lane layouts (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
s = bk.HMCDiag(bk.IsoGauss(100), 0.1, 10, chains=1003, seed=0); s.sample_n(3)
s = bk.HMCDiag(bk.DiagGauss(np.zeros(77), np.ones(77) * 2), 0.1, 4, chains=515, seed=0) if hasattr(bk, "DiagGauss") else s; s.sample_n(2)
s = bk.MALA(bk.IsoGauss(100), 0.05, chains=1003, seed=0); s.sample_n(3)
s = bk.DrGhmcDiag(bk.IsoGauss(100), 2, [0.6, 0.2], [5, 10], 0.5, chains=517, seed=0); s.sample_n(2)
D, M, T = 50, 5001, 3
mu = np.random.default_rng(0).normal(size=D)
model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
smc = bk.TemperedLikelihoodSMC(model, M, T, torch.randn(M, D, device="cuda"), bk.metropolis_kernel(0.2), resample="systematic", seed=1)
smc.run()
torch.cuda.synchronize()
print("ok", float(smc.thetas.mean()))
