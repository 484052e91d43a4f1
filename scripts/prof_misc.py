"""Small profiling drivers: python scripts/prof_misc.py {c1|smc|c3|ess|acf|sort}"""
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
elif w == "c3":
    from oracle.models import HierLogReg
    N, Dx, C = 100_000, 100, 1024
    X, y = HierLogReg.c3_data(N, Dx, seed=0)
    s = bk.HMCDiag(bk.HierLogReg(X, y), 0.01, 10, init=np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1, seed=0)
    s.sample_n(3)
elif w == "sort":
    x = torch.randn(10000, 2048, 1, device="cuda")
    bk.rank_normalized_rhat(x, draws_first=True); bk.rank_normalized_rhat(x, draws_first=True)
elif w in ("ess", "acf"):
    N, S = 10000, 25600 if w == "ess" else 4096
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    phi = torch.rand(S, device="cuda", generator=g) * 0.9
    x = torch.empty(N, S, device="cuda"); cur = torch.randn(S, device="cuda", generator=g)
    for t in range(N):
        cur = phi * cur + torch.randn(S, device="cuda", generator=g); x[t] = cur
    xs = x.t().contiguous()
    if w == "ess":
        bk.ess(xs); bk.ess(xs)
    else:
        bk.autocorr(xs); bk.autocorr(xs)
torch.cuda.synchronize()
print("done")
