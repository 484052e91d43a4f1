"""GPU diagnostic: tensor-core HMC path vs the fp64 oracle, and kernel timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib
from oracle import samplers as osm
from oracle.models import DensePrecGauss

def np_(t): return t.detach().cpu().numpy().astype(np.float64)

for (D, C, L, eps) in [(1000, 300, 1, 0.2), (1000, 300, 4, 0.1), (200, 513, 10, 0.1), (128, 64, 3, 0.2)]:
    rng = np.random.default_rng(L * 7 + D)
    P = DensePrecGauss.c2_precision(D, 0)
    th0 = rng.normal(size=(C, D)).astype(np.float32).astype(np.float64)
    n = 2
    zs = rng.standard_normal((n, C, D)).astype(np.float32).astype(np.float64)
    us = rng.random((n, C)).astype(np.float32).astype(np.float64)
    od, ol, oa = osm.hmc_diag_batch(DensePrecGauss(P), th0, zs, us, eps, L)
    for tc_off in ("0", "1"):
        os.environ["BK_DISABLE_TC"] = tc_off
        s = bk.HMCDiag(bk.DensePrecGauss(P, dtype=torch.float32), eps, L, init=th0)
        d, l = s.sample_n(n, normals=zs, uniforms=us)
        d, l, a = np_(d), np_(l), np_(s.last_accept).astype(bool)
        for t in range(n):
            flips = (a[t] != oa[t])
            ok = ~(a[:t + 1] != oa[:t + 1]).any(0)
            print(f"D={D} C={C} L={L} tc_off={tc_off} draw{t}: flips={flips.sum()} "
                  f"max|dtheta|={np.abs(d[t][ok]-od[t][ok]).max():.3e} max|dlogp|={np.abs(l[t][ok]-ol[t][ok]).max():.3e} "
                  f"acc={a[t].mean():.3f}/{oa[t].mean():.3f}", flush=True)
os.environ["BK_DISABLE_TC"] = "0"

# timings at bench scale
lib = _lib.lib()
D, C = 1000, 65536
P = DensePrecGauss.c2_precision(D, 0)
model = bk.DensePrecGauss(P, dtype=torch.float32)
for L in (1, 2, 10):
    s = bk.HMCDiag(model, 0.1, L, chains=C, seed=0)
    s.sample_n(2)
    torch.cuda.synchronize()
    lib.bk_profile_enable(1)
    t0 = time.perf_counter()
    s.sample_n(5)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    ms, nn = _lib.f64(0), _lib.u64(0)
    lib.bk_profile_read(_lib.PROF_GRAD, ms, nn)
    lib.bk_profile_enable(0)
    print(f"L={L}: {dt*1e3:.3f} ms/draw; gemm launches {nn.value} total {ms.value/5:.3f} ms/draw "
          f"avg {ms.value/max(nn.value,1):.3f} ms", flush=True)
