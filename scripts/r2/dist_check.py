"""Multi-GPU checks on real hardware (one process per GPU, NCCL + NVLink peer memory):
   torchrun --nproc-per-node N scripts/r2/dist_check.py [--c4]
 (1) sharded TemperedLikelihoodSMC (systematic + multinomial, injected streams) == oracle, bit-exact indices
 (2) fp32 Philox SMC: checksum of the final particles is identical for every N (printed; compare across runs)
 (3) cross-rank R-hat (moment all-reduce) == oracle
 (4) sharded Stretcher (complementary-half all-gather) == oracle
 (5) --c4: BASELINE config c4 (1M particles x 50 dims x 100 temperatures, STRONG scaling: M is global)
     timed with CUDA events, max over ranks; bytes on the wire per temperature.
Rank 0 prints one JSON line per item."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bayes_kit_b200 as bk
from oracle import diagnostics as od
from oracle import samplers as osm
from oracle.models import DiagGauss, GaussPriorLik

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
np_ = lambda t: t.detach().cpu().numpy()


def say(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def gather_np(a):
    if world == 1:
        return a
    out = [None] * world
    dist.all_gather_object(out, a)
    return np.concatenate(out)


# (1) SMC parity
rng = np.random.default_rng(5)
D, M, T, scale = 6, 1031, 6, 0.25
mu = rng.normal(size=D)
om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
th0 = rng.normal(size=(M, D))
zs, au, ru = rng.standard_normal((T, M, D)), rng.random((T, M)), rng.random((T, M))
model64 = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64)
for mode in ("systematic", "multinomial"):
    oth, oidx = osm.smc_tempered(om, th0, zs, au, ru, scale, T, resample=mode)
    smc = bk.TemperedLikelihoodSMC(model64, M, T, th0, bk.metropolis_kernel(scale), resample=mode)
    lo, hi = smc._lo, smc._hi
    ok = True
    for n in range(1, T + 1):
        smc.transition(n, normals=zs[n - 1, lo:hi], acc_uniforms=au[n - 1, lo:hi],
                       res_uniforms=ru[n - 1, :1] if mode == "systematic" else ru[n - 1, lo:hi])
        th_n = smc.thetas            # waits for every rank's indices of this step
        ok = ok and np.array_equal(np_(smc.last_indices), oidx[n - 1][lo:hi])
    ok = ok and np.allclose(np_(smc.thetas), oth[lo:hi], rtol=1e-12, atol=1e-12)
    oks = gather_np(np.array([ok]))
    say(check="smc_parity_vs_oracle", mode=mode, world=world, ok=bool(oks.all()), ess=smc.weight_ess[:3])
    assert oks.all(), (mode, rank)

# (2) Philox invariance: checksum of the final particles
D2, M2, T2 = 50, 200_000, 20
g = torch.Generator(device="cuda").manual_seed(1)
mu2 = torch.randn(D2, device="cuda", generator=g)
th02 = torch.randn(M2, D2, device="cuda", generator=g)     # same on every rank (same generator seed)
model32 = bk.GaussPriorLik(torch.zeros(D2), torch.ones(D2), mu2, 4 * torch.ones(D2))
for kern, name in ((bk.metropolis_kernel(0.2), "rw"), (bk.mala_kernel(0.02), "mala"), (bk.hmc_kernel(0.1, 3), "hmc")):
    smc = bk.TemperedLikelihoodSMC(model32, M2, T2, th02, kern, resample="systematic", seed=7)
    smc.run()
    th = smc.thetas.double()
    # exact, order-independent integer checksum: bit patterns weighted by a function of the GLOBAL row id
    bits = smc.thetas.view(torch.int32).to(torch.int64)
    rows = torch.arange(smc._lo, smc._hi, device="cuda").reshape(-1, 1)
    h = torch.stack([bits.sum(), (bits * ((rows * 31 + torch.arange(D2, device="cuda")) % 1009 + 1)).sum()])
    if world > 1:
        dist.all_reduce(h)
    say(check="smc_philox_checksum", kernel=name, world=world, M=M2, T=T2, checksum=[int(v) for v in h.tolist()],
        mean_err=float((th.mean(0) - 0.8 * mu2.double()).abs().max()) if name == "rw" else None)

# (3) R-hat
rng = np.random.default_rng(3)
C, N, P = 64, 500, 7
x = rng.normal(size=(C, N, P)) + rng.normal(size=(C, 1, P)) * 0.2 + 1e6
want = np.array([od.rhat(list(x[:, :, p])) for p in range(P)])
lo, hi = bk.dist.shard_range(C, rank, world)
got = np_(bk.rhat(torch.as_tensor(x[lo:hi], device="cuda"), group=dist.group.WORLD if world > 1 else None))
err = float(np.max(np.abs(got / want - 1)))
say(check="rhat_allreduce_vs_oracle", world=world, max_rel_err=err, wire_bytes_per_rank=P * 4 * 8)
assert err < 1e-9

# (4) Stretcher
rng = np.random.default_rng(9)
D, W, n = 7, 64, 5
mu_, pr_ = rng.normal(size=D), rng.uniform(0.5, 3, D)
th0 = rng.normal(size=(W, D))
us = rng.random((n, W, 3))
want, wacc = osm.stretch(DiagGauss(mu_, pr_), th0, us, a=2.0)
s = bk.Stretcher(bk.DiagGauss(mu_, pr_, dtype=torch.float64), a=2.0, walkers=W, init=th0)
h = W // 2
ok = True
for t in range(n):
    u_loc = np.concatenate([us[t, s._lo:s._hi], us[t, h + s._lo:h + s._hi]])
    got = np_(s.sample(uniforms=u_loc))
    mine = np.concatenate([want[t][s._lo:s._hi], want[t][h + s._lo:h + s._hi]])
    ok = ok and np.allclose(got, mine, rtol=1e-10, atol=1e-10)
oks = gather_np(np.array([ok]))
say(check="stretcher_allgather_vs_oracle", world=world, ok=bool(oks.all()))
assert oks.all()

# (5) c4
if "--c4" in sys.argv:
    D4, M4, T4 = 50, 1_000_000, 100
    g = torch.Generator(device="cuda").manual_seed(0)
    mu4 = torch.randn(D4, device="cuda", generator=g)
    model4 = bk.GaussPriorLik(torch.zeros(D4), torch.ones(D4), mu4, 4 * torch.ones(D4))
    lo, hi = bk.dist.shard_range(M4, rank, world)
    th04 = torch.randn(M4, D4, device="cuda", generator=g)[lo:hi].contiguous()
    for rep in range(3):
        smc = bk.TemperedLikelihoodSMC(model4, M4, T4, th04, bk.metropolis_kernel(0.2), resample="systematic", seed=1)
        smc.transition(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        smc.run_steps(2, T4)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        smc._check()
        th = smc.thetas.double()
        m1 = th.sum(0)
        if world > 1:
            dist.all_reduce(m1)
        ms = float(ms)
        say(check="c4_smc_strong_scaling", world=world, M=M4, D=D4, T=T4, ms_total_99=ms, ms_per_temperature=ms / (T4 - 1),
            particle_steps_per_s=M4 * (T4 - 1) / (ms * 1e-3), wire_bytes_per_rank_per_temperature=smc.wire_bytes_per_step(),
            posterior_mean_max_err=float((m1 / M4 - 0.8 * mu4.double()).abs().max()), ess_last=smc.weight_ess[-1])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
