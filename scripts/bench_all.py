"""Secondary workloads (BASELINE configs c1, c4, c5): device-resident timings
with CUDA events; one JSON line per workload (for profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
HBM = 6548.5  # GB/s measured (MEASURED_PEAKS.json)
which = sys.argv[1:] or ["c1", "c1mala", "c2mala", "c3", "drghmc", "c4", "c5", "acf", "rnrhat", "stretch"]

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if "c1" in which:
    D, L, n = 100, 10, 10
    for C in (65536, 1048576):
        s = bk.HMCDiag(bk.IsoGauss(D), 0.1, L, chains=C, seed=0)
        ms = timed(lambda: s.sample_n(n))
        steps = C * n / (ms * 1e-3)
        print(json.dumps({"workload": f"c1 HMCDiag iso D={D} L={L} C={C} n={n}/launch fp32", "ms_per_launch": ms,
              "chain_steps_per_s": steps, "GBps_algorithmic(804B/step)": steps * 804 / 1e9,
              "hbm_frac": steps * 804 / 1e9 / HBM, "flops_per_step~(5L+8)D": (5*L+8)*D,
              "TFLOPs": steps * (5*L+8)*D / 1e12}), flush=True)
if "c1mala" in which:
    D, n, C = 100, 10, 1048576
    s = bk.MALA(bk.IsoGauss(D), 0.05, chains=C, seed=0)
    ms = timed(lambda: s.sample_n(n))
    steps = C * n / (ms * 1e-3)
    print(json.dumps({"workload": f"MALA iso D={D} C={C} n={n}/launch fp32", "ms_per_launch": ms,
          "chain_steps_per_s": steps, "GBps_algorithmic(804B/step)": steps * 804 / 1e9,
          "hbm_frac": steps * 804 / 1e9 / HBM}), flush=True)
if "c2mala" in which:
    from oracle.models import DensePrecGauss
    D, C = 1000, 65536
    model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0))
    s = bk.MALA(model, 2e-3, chains=C, seed=0)
    ms = timed(lambda: s.sample_n(1), reps=10, warm=3)
    print(json.dumps({"workload": f"c2 MALA dense D={D} C={C} eps=2e-3 fp32 (begin -> tcgen05 3-pass gradient -> accept)",
          "ms_per_draw": ms, "chain_steps_per_s": C / (ms * 1e-3), "accept": float(s.last_accept.float().mean()),
          "grad_TFLOPs(2CD^2)": 2.0 * C * D * D / (ms * 1e-3) / 1e12}), flush=True)
if "c3" in which:
    from oracle.models import HierLogReg
    N, Dx, C = 100_000, 100, 1024
    X, y = HierLogReg.c3_data(N, Dx, seed=0)
    model = bk.HierLogReg(X, y)
    th0 = np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1
    s = bk.HMCDiag(model, 0.01, 10, init=th0, seed=0)
    ms = timed(lambda: s.sample_n(1), reps=3, warm=1)
    print(json.dumps({"workload": f"c3 HMCDiag hier-logreg N={N} Dx={Dx} C={C}/GPU L=10 eps=0.01 fp32 (tcgen05: bf16 gradients at every step, split-precision density at the endpoint)",
          "ms_per_draw": ms, "chain_steps_per_s": C / (ms * 1e-3), "accept": float(s.last_accept.float().mean()),
          "grad_evals_per_s": 10 * C / (ms * 1e-3), "TFLOPs(4NDx per grad)": 10 * C * 4.0 * N * Dx / (ms * 1e-3) / 1e12}), flush=True)
if "drghmc" in which:
    D, C, n = 100, 262144, 10
    s = bk.DrGhmcDiag(bk.IsoGauss(D), 2, [0.6, 0.2], [5, 10], 0.5, chains=C, seed=0)
    ms = timed(lambda: s.sample_n(n))
    print(json.dumps({"workload": f"DrGhmcDiag iso D={D} K=2 eps=(0.6,0.2) counts=(5,10) damping=0.5 C={C} n={n}/launch fp32",
          "ms_per_launch": ms, "chain_steps_per_s": C * n / (ms * 1e-3),
          "move_rate": float(s.last_accept.float().mean())}), flush=True)
if "c4" in which:
    D, M, T = 50, 1_000_000, 100
    mu = np.random.default_rng(0).normal(size=D)
    model = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
    th0 = torch.randn(M, D, device="cuda")
    for mode in ("systematic", "multinomial"):
        smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample=mode, seed=1)
        smc.transition(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for n_ in range(2, T + 1): smc.transition(n_)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rate = M * (T - 1) / dt
        th = smc.thetas.double()
        print(json.dumps({"workload": f"c4 SMC M={M} D={D} T={T} {mode} fp32", "s_total": dt,
              "particle_steps_per_s": rate, "GBps_algorithmic(816B)": rate * 816 / 1e9,
              "hbm_frac": rate * 816 / 1e9 / HBM, "post_mean_err_max": float((th.mean(0).cpu() - torch.tensor(0.8 * mu)).abs().max()),
              "post_var_mean": float(th.var(0).mean())}), flush=True)
if "c5" in which:
    N, P = 10000, 100
    for Cn, layout in ((256, "series-major [chains,params,draws]"), (256, "draws-first [draws,chains,params]")):
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        phi = torch.rand(Cn, P, device="cuda", generator=g) * 0.9
        x = torch.empty(N, Cn, P, device="cuda")
        cur = torch.randn(Cn, P, device="cuda", generator=g)
        for t in range(N):
            cur = phi * cur + torch.randn(Cn, P, device="cuda", generator=g)
            x[t] = cur
        if layout.startswith("series"):
            xs = x.permute(1, 2, 0).contiguous().reshape(Cn * P, N)   # [series, draws]
            f_ess = lambda: bk.ess(xs); f_rhat = None
        else:
            xs = x
            f_ess = lambda: bk.ess(xs, draws_first=True); f_rhat = lambda: bk.rhat(xs, draws_first=True)
        ms = timed(f_ess, reps=3, warm=1)
        ser = Cn * P / (ms * 1e-3)
        e = f_ess()
        iat_true = ((1 + phi) / (1 - phi))
        rel = float(((N / e.reshape(Cn, P)) / iat_true - 1).abs().median())
        out = {"workload": f"c5 ess N={N} series={Cn*P} {layout} fp32 in / fp64 acc", "ms": ms, "series_per_s": ser,
               "GBps_algorithmic(N*4B)": ser * N * 4 / 1e9, "hbm_frac": ser * N * 4 / 1e9 / HBM,
               "median_rel_err_vs_AR1_closed_form": rel}
        if f_rhat:
            ms2 = timed(f_rhat, reps=3, warm=1)
            out.update({"rhat_ms": ms2, "rhat_GBps": Cn * P * N * 4 / 1e9 / (ms2 * 1e-3)})
        print(json.dumps(out), flush=True)

if "acf" in which:
    N, S_ = 10000, 8192
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    phi = torch.rand(S_, device="cuda", generator=g) * 0.9
    x = torch.empty(N, S_, device="cuda"); cur = torch.randn(S_, device="cuda", generator=g)
    for t in range(N):
        cur = phi * cur + torch.randn(S_, device="cuda", generator=g); x[t] = cur
    xs = x.t().contiguous()                      # [series, draws]
    ms = timed(lambda: bk.autocorr(xs), reps=3, warm=1)
    a = bk.autocorr(xs[:64])
    ref = torch.stack([phi[:64] ** k for k in range(4)], 1)      # AR(1): rho_k = phi^k
    print(json.dumps({"workload": f"c5 autocorr (all {N} lags, on-chip real FFT of length 20480) series={S_} fp32 in / fp64", "ms": ms,
          "series_per_s": S_ / (ms * 1e-3), "GBps_algorithmic(N*(4+8)B)": S_ * N * 12 / 1e9 / (ms * 1e-3),
          "max_abs_err_first_lags_vs_AR1": float((a[:, :4] - ref.double()).abs().max())}), flush=True)

if "rnrhat" in which:
    N, Cn, P = 10000, 4096, 2
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn(N, Cn, P, device="cuda", generator=g)          # [draws, chains, params], the samplers' layout
    ms = timed(lambda: bk.rank_normalized_rhat(x, draws_first=True), reps=3, warm=1)
    ms_s = timed(lambda: bk.split_rhat(x, draws_first=True), reps=3, warm=1)
    r = bk.rank_normalized_rhat(x, draws_first=True)
    n = N * Cn * P
    print(json.dumps({"workload": f"rank_normalized_rhat {Cn} chains x {N} draws x {P} params fp32 (radix sort of {N*Cn/1e6:.0f} M keys per parameter)",
          "ms": ms, "M_draws_per_s": n / (ms * 1e-3) / 1e6, "GBps_sort_traffic(80B/key)": n * 80 / 1e9 / (ms * 1e-3),
          "split_rhat_ms": ms_s, "split_rhat_GBps": n * 4 / 1e9 / (ms_s * 1e-3), "value": [float(v) for v in r]}), flush=True)

if "stretch" in which:
    from oracle.models import DensePrecGauss
    D, W = 1000, 65536
    s = bk.Stretcher(bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0)), walkers=W, seed=0)
    ms = timed(lambda: s.sample(), reps=10, warm=3)
    print(json.dumps({"workload": f"Stretcher (stretch move) {W} walkers x {D}-dim dense Gaussian fp32, one sweep",
          "ms_per_sweep": ms, "walker_moves_per_s": W / (ms * 1e-3), "accept": float(s.last_accept.float().mean())}), flush=True)
