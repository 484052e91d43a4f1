"""c4 timing on one GPU (CUDA events over temperatures 2..T): python scripts/r2/time_smc.py [M] [D] [T]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 50
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device="cuda").manual_seed(0)
mu = torch.randn(D, device="cuda", generator=g)
model = bk.GaussPriorLik(torch.zeros(D), torch.ones(D), mu, 4 * torch.ones(D))
th0 = torch.randn(M, D, device="cuda", generator=g)
for mode in ("systematic", "multinomial"):
    best = 1e9
    for rep in range(3):
        smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample=mode, seed=1)
        smc.transition(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        smc.run_steps(2, T)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    chk = float(smc.thetas.double().sum())
    print(json.dumps({"M": M, "D": D, "T": T, "mode": mode, "layout": os.environ.get("BK_SMC_LAYOUT", "default"),
                      "checksum": chk,
                      "ms_total": best, "ms_per_temperature": best / (T - 1), "particle_steps_per_s": M * (T - 1) / best * 1e3}), flush=True)
