"""c2 pipeline: end + begin fused into one 'turn' kernel between the draws of a sample_n call (default) vs the two
separate kernels through theta / grad (BK_TC_TURN=0): same arithmetic, so draws / log densities / final state must
agree bit for bit.  Also prints ms per draw.  python scripts/r2/turn_check.py  (spawns itself twice per shape)."""
import hashlib, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) > 1 and sys.argv[1] != "quick":
    import torch
    import bayes_kit_b200 as bk
    from oracle.models import DensePrecGauss
    C, D, L, n = (int(v) for v in sys.argv[1:5])
    model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
    s = bk.HMCDiag(model, 0.1, L, chains=C, seed=3) if L > 0 else bk.MALA(model, 2e-3, chains=C, seed=3)
    d, lp = s.sample_n(n)
    d2, lp2 = s.sample_n(2)          # second call: starts from the state the first one left in theta / grad
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample_n(8); e1.record(); torch.cuda.synchronize()
    hh = lambda x: hashlib.sha1(x.cpu().numpy().tobytes()).hexdigest()[:12]
    print(hh(d), hh(lp), hh(d2), hh(lp2), hh(s.theta), f"{float(s.last_accept.float().mean()):.4f}",
          f"ms_per_draw={e0.elapsed_time(e1) / 8:.4f}")
else:
    ok = True
    quick = len(sys.argv) > 1
    for shape in ((65536, 1000, 10, 5), (1000, 250, 7, 4), (4096, 1000, 1, 4), (70000, 1000, 3, 3)) if quick else ((65536, 1000, 10, 5), (4096, 1000, 10, 4), (1000, 250, 7, 4), (3000, 1001, 4, 3), (70000, 1000, 3, 3),
                  (512, 2000, 12, 3), (4096, 1000, 1, 4), (4096, 1000, 2, 4), (65536, 1000, 0, 5)):
        outs = []
        for turn in ("1", "0"):
            env = dict(os.environ, BK_TC_TURN=turn)
            r = subprocess.run(["timeout", "120", sys.executable, __file__, *map(str, shape)], env=env, capture_output=True, text=True)
            outs.append(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else f"FAILED rc={r.returncode} {r.stderr[-300:]}")
        same = outs[0].split(" ms_per_draw")[0] == outs[1].split(" ms_per_draw")[0] and not outs[0].startswith("FAILED")
        ok &= same
        print(shape, "IDENTICAL" if same else "DIFFERENT", outs, flush=True)
    print("TURN_CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
