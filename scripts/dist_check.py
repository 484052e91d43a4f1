"""torchrun --nproc-per-node N scripts/dist_check.py : multi-GPU checks over NCCL.
 (1) sharded HMC == the same chains of a single-GPU run (Philox keyed by global chain id);
 (2) sharded TemperedLikelihoodSMC (multinomial, injected streams) reproduces the oracle's
     resample indices / particles exactly -- all-gather of log-weights + particles;
 (3) cross-chain R-hat from sharded chains == oracle on all chains;
 (4) sharded Stretcher (complementary-walker all-gather) == the oracle's single ensemble."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bayes_kit_b200 as bk
from bayes_kit_b200 import dist as bd
from oracle import samplers as osm, diagnostics as od
from oracle.models import GaussPriorLik
dev = torch.device("cuda", local)
np_ = lambda t: t.detach().cpu().numpy()

# (1) HMC shard invariance
model = bk.IsoGauss(20, device=dev)
init = np.random.default_rng(1).normal(size=(64, 20))
lo, hi = bd.shard_range(64, rank, world)
part = bk.HMCDiag(model, 0.2, 5, init=init[lo:hi], seed=9, chain_offset=lo).sample_n(10)[0]
full = bk.HMCDiag(model, 0.2, 5, init=init, seed=9).sample_n(10)[0]
assert torch.equal(full[:, lo:hi], part), "shard invariance"

# (2) SMC parity, sharded
rng = np.random.default_rng(3)
D, M, T, scale = 6, 203, 5, 0.25
mu = rng.normal(size=D)
om = GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D))
th0 = rng.normal(size=(M, D)); zs = rng.standard_normal((T, M, D)); au = rng.random((T, M)); ru = rng.random((T, M))
for mode in ("multinomial", "systematic"):
    oth, oidx = osm.smc_tempered(om, th0, zs, au, ru, scale, T, resample=mode)
    m = bk.GaussPriorLik(np.zeros(D), np.ones(D), mu, 4 * np.ones(D), dtype=torch.float64, device=dev)
    smc = bk.TemperedLikelihoodSMC(m, M, T, th0, bk.metropolis_kernel(scale), resample=mode)
    lo, hi = bd.shard_range(M, rank, world)
    for n in range(1, T + 1):
        smc.transition(n, normals=zs[n - 1, lo:hi], acc_uniforms=au[n - 1, lo:hi],
                       res_uniforms=ru[n - 1] if mode == "multinomial" else ru[n - 1, :1])
        assert np.array_equal(np_(smc.last_indices), oidx[n - 1][lo:hi]), (mode, n)
    assert np.allclose(np_(smc.thetas), oth[lo:hi], rtol=1e-12, atol=1e-12)

# (3) R-hat over sharded chains
ch = rng.normal(size=(12, 200, 3)) + rng.normal(size=(12, 1, 3))
lo, hi = bd.shard_range(12, rank, world)
r = bk.rhat(torch.as_tensor(ch[lo:hi], device=dev))
assert np.allclose(np_(r), od.rhat_batch(ch), rtol=1e-12)
# (4) Stretcher: walkers of both halves sharded, complementary half all-gathered per half-step
D, W, n = 5, 24, 4
om4 = __import__("oracle.models", fromlist=["IsoGauss"]).IsoGauss(D, 1.3)
th0 = rng.normal(size=(W, D)); us = rng.random((n, W, 3))
want, wacc = osm.stretch(om4, th0, us, a=2.0)
st_ = bk.Stretcher(bk.IsoGauss(D, 1.3, dtype=torch.float64, device=dev), a=2.0, walkers=W, init=th0)
h = W // 2
lo, hi = bd.shard_range(h, rank, world)
rows = np.r_[lo:hi, h + lo:h + hi]                      # this rank's walkers: its slice of each half
for t in range(n):
    got = st_.sample(uniforms=us[t][rows])
    assert np.allclose(np_(got), want[t][rows], rtol=1e-10, atol=1e-10), ("stretcher", t)
    assert np.array_equal(np_(st_.last_accept).astype(bool), wacc[t][rows])
dist.barrier()
if rank == 0:
    print(f"dist_check ok on {world} GPUs")
dist.destroy_process_group()
