"""Secondary legs of bench.py: BASELINE configs c1, c3, c4, c5 and MALA on c2, each as
{"workload", "value", "unit", "ms", "roofline": {"bound", "achieved", "peak", "unit", "frac"}}.
Device-resident timings with CUDA events after warm-up; multi-rank legs take the max over ranks.
(bench.py's headline stays c2 HMCDiag; these lines go under config.secondary.)"""
import numpy as np
import torch
import torch.distributed as dist

import bayes_kit_b200 as bk


def _timed(fn, reps=5, warm=2, median=False):
    """ms per call on the current stream; median=True times every call separately and returns the median (calls that
    allocate hundreds of MB per invocation see an occasional caching-allocator stall that is not kernel time)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if median:
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _max_over_ranks(ms, world):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def _roof(bound, achieved, peak, unit):
    return {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak}


def leg_c1(hbm):
    """c1 shape on the GPU: HMCDiag, 100-dim std normal, eps 0.1, L 10, 1,048,576 chains, 10 draws per launch.
    Algorithmic bytes 804 B per chain-step."""
    D, L, n, C = 100, 10, 10, 1048576
    s = bk.HMCDiag(bk.IsoGauss(D), 0.1, L, chains=C, seed=0)
    ms = _timed(lambda: s.sample_n(n), reps=5, warm=2)
    v = C * n / (ms * 1e-3)
    return {"workload": f"c1: HMCDiag std normal D={D} L={L} eps=0.1, {C} chains x {n} draws per launch, fp32 Philox",
            "value": v, "unit": "chain-steps/s", "ms": ms, "roofline": _roof("hbm", v * 804 / 1e9, hbm, "GB/s")}


def leg_c2_mala(hbm, tf):
    """MALA on the c2 model (BASELINE configs[1] names MALA + HMCDiag): one 3-pass tcgen05 gradient per draw."""
    from bench import c2_precision
    D, C = 1000, 65536
    s = bk.MALA(bk.DensePrecGauss(c2_precision()), 2e-3, chains=C, seed=0)
    ms = _timed(lambda: s.sample_n(1), reps=10, warm=3)
    v = C / (ms * 1e-3)
    # a MALA draw = begin (18 B per element) + one 3-pass tcgen05 gradient (10 B) + end (24 B): the row kernels'
    # HBM streams bound it, the GEMM is 0.3 of the 1.07 ms
    return {"workload": f"c2: MALA eps=2e-3, {C} chains x {D}-dim dense-precision Gaussian, fp32 Philox", "value": v,
            "unit": "chain-steps/s", "ms": ms, "accept_rate": float(s.last_accept.float().mean()),
            "roofline": _roof("hbm", 52.0 * C * D / (ms * 1e-3) / 1e9, hbm, "GB/s"),
            "tensor_side": _roof("tensor", 3 * 2.0 * C * D * D / (ms * 1e-3) / 1e12, tf, "TFLOP/s")}


def leg_c3(rank, world, tf):
    """c3: HMCDiag on hierarchical logistic regression, N=100k x D=100, 1024 chains per GPU (8,192 over 8 GPUs),
    X replicated per GPU, no communication.  Tensor-bound: 4 N Dx flop per gradient per chain."""
    N, Dx, C, L = 100_000, 100, 1024, 10
    rng = np.random.default_rng(0)                # X ~ N(0,1)/sqrt(D), y ~ Bernoulli(sigmoid(X beta*))
    X = rng.normal(size=(N, Dx)) / np.sqrt(Dx)
    beta = rng.normal(size=Dx)
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
    model = bk.HierLogReg(X, y)
    th0 = np.random.default_rng(1 + rank).normal(size=(C, Dx + 2)) * 0.1
    s = bk.HMCDiag(model, 0.01, L, init=th0, seed=0, chain_offset=rank * C)
    ms = _max_over_ranks(_timed(lambda: s.sample_n(1), reps=5, warm=2), world)
    v = world * C / (ms * 1e-3)
    return {"workload": f"c3: HMCDiag hier-logreg N={N} Dx={Dx}, {C} chains/GPU x {world} GPU, L={L} eps=0.01, fp32",
            "value": v, "unit": "chain-steps/s", "ms": ms, "accept_rate": float(s.last_accept.float().mean()),
            "grad_evals_per_s": v * L,
            "roofline": _roof("tensor", L * C * 4.0 * N * Dx / (ms * 1e-3) / 1e12, tf, "TFLOP/s")}


def leg_c4(rank, world, hbm, weak=False):
    """c4: TemperedLikelihoodSMC, 50-dim, 100 temperatures, systematic resampling, RW-Metropolis kernel.
    strong: 1M particles in total, sharded over the ranks (BASELINE configs[3]); weak: 1M per GPU.
    Algorithmic bytes 816 B per particle-temperature.  No host-side collective per
    temperature: mailbox messages + peer index stores + peer row reads inside the kernels."""
    D, T = 50, 100
    M = 1_000_000 * (world if weak else 1)
    g = torch.Generator(device="cuda").manual_seed(0)
    mu = torch.randn(D, device="cuda", generator=g)
    model = bk.GaussPriorLik(torch.zeros(D), torch.ones(D), mu, 4 * torch.ones(D))
    lo, hi = bk.dist.shard_range(M, rank, world)
    g2 = torch.Generator(device="cuda").manual_seed(100 + rank)
    th0 = torch.randn(hi - lo, D, device="cuda", generator=g2)
    best = None
    for rep in range(3):
        smc = bk.TemperedLikelihoodSMC(model, M, T, th0, bk.metropolis_kernel(0.2), resample="systematic", seed=1)
        smc.transition(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smc.run_steps(2, T)               # 99 temperatures enqueued by one library call
        e1.record()
        torch.cuda.synchronize()
        ms = _max_over_ranks(e0.elapsed_time(e1), world)
        smc._check()
        best = ms if best is None else min(best, ms)
        wire = smc.wire_bytes_per_step()
        th = smc.thetas.double().sum(0)
        if world > 1:
            dist.all_reduce(th)
            dist.barrier()
        err = float((th / M - 0.8 * mu.double()).abs().max())
    v = M * (T - 1) / (best * 1e-3)
    return {"workload": f"c4 ({'weak' if weak else 'strong'} scaling): TemperedLikelihoodSMC M={M} D={D} T={T} systematic, "
                        f"{world} GPU, fp32 Philox", "value": v, "unit": "particle-steps/s", "ms": best,
            "ms_per_temperature": best / (T - 1), "launches_per_temperature": 3,
            "wire_bytes_per_rank_per_temperature": wire, "host_collectives_per_temperature": 0,
            "posterior_mean_max_err": err,
            "roofline": _roof("hbm", v * 816 / 1e9 / world, hbm, "GB/s")}


def leg_c5(hbm):
    """c5 at reduced chain count (256 chains x 10k draws x 100 params = 25,600 series; the full 819,200 series run
    is tests/test_gpu_full_size.py): ess (Geyer IMSE), all-lag autocorr (hand-written FFT), rhat.
    Algorithmic bytes: one read of every draw (N * 4 B per series; + N * 8 B written for autocorr)."""
    N, P, Cn = 10000, 100, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    phi = torch.rand(Cn, P, device="cuda", generator=g) * 0.9
    x = torch.empty(N, Cn, P, device="cuda")
    cur = torch.randn(Cn, P, device="cuda", generator=g)
    for t in range(N):
        cur = phi * cur + torch.randn(Cn, P, device="cuda", generator=g)
        x[t] = cur
    xs = x.permute(1, 2, 0).contiguous().reshape(Cn * P, N)      # [series, draws]
    ms_e = _timed(lambda: bk.ess(xs), reps=3, warm=1)
    ms_r = _timed(lambda: bk.rhat(x, draws_first=True), reps=3, warm=1)
    xa = xs[:8192]
    ms_a = _timed(lambda: bk.autocorr(xa), reps=7, warm=2, median=True)
    S = Cn * P
    return {"workload": f"c5 (reduced): {Cn} chains x {N} draws x {P} params fp32 in / fp64 accumulate",
            "ess": {"value": S / (ms_e * 1e-3), "unit": "series/s", "ms": ms_e,
                    "roofline": _roof("hbm", S * N * 4 / 1e9 / (ms_e * 1e-3), hbm, "GB/s")},
            "rhat": {"value": S / (ms_r * 1e-3), "unit": "series/s", "ms": ms_r,
                     "roofline": _roof("hbm", S * N * 4 / 1e9 / (ms_r * 1e-3), hbm, "GB/s")},
            "autocorr": {"value": 8192 / (ms_a * 1e-3), "unit": "series/s", "ms": ms_a,
                         "roofline": _roof("hbm", 8192 * N * 12 / 1e9 / (ms_a * 1e-3), hbm, "GB/s")}}
