"""Small profiling drivers: python scripts/prof_misc.py {c1|smc|ess|acf}"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
w = sys.argv[1]
if w == "c1":
    s = bk.HMCDiag(bk.IsoGauss(100), 0.1, 10, chains=1048576, seed=0)
    s.sample_n(10); s.sample_n(10)
elif w == "smc":
    D, M, T = 50, 1_000_000, 100
    mu = np.random.default_rng(0).normal(size=D)
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    smc = bk.TemperedLikelihoodSMC(model, M, T, torch.randn(M, D, device="cuda"), bk.metropolis_kernel(0.2),
                                   resample="systematic", seed=1)
    for n in range(1, 5): smc.transition(n)
elif w in ("ess", "acf"):
    N, S = 10000, 4096
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    phi = torch.rand(S, device="cuda", generator=g) * 0.9
    x = torch.empty(N, S, device="cuda"); cur = torch.randn(S, device="cuda", generator=g)
    for t in range(N):
        cur = phi * cur + torch.randn(S, device="cuda", generator=g); x[t] = cur
    xs = x.t().contiguous()
    if w == "ess":
        bk.ess(xs); bk.ess(xs)
    else:
        bk.autocorr(xs[:512]); bk.autocorr(xs[:512])
torch.cuda.synchronize()
print("done")
