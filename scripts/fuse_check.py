"""Fused multi-step STEP launch vs one launch per leapfrog step (BK_TC_FUSE=0): the arithmetic is identical, so the
draws must agree bit for bit.  python scripts/fuse_check.py  (spawns itself twice per shape)."""
import hashlib, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import bayes_kit_b200 as bk
    from oracle.models import DensePrecGauss
    C, D, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    model = bk.DensePrecGauss(DensePrecGauss.c2_precision(D, 0), dtype=torch.float32)
    s = bk.HMCDiag(model, 0.1, L, chains=C, seed=3)
    d, lp = s.sample_n(3)
    torch.cuda.synchronize()
    print(hashlib.sha1(d.cpu().numpy().tobytes()).hexdigest(), hashlib.sha1(lp.cpu().numpy().tobytes()).hexdigest(),
          float(s.last_accept.float().mean()))
else:
    ok = True
    for shape in ((4096, 1000, 10), (65536, 1000, 10), (1000, 250, 7), (70000, 1000, 3), (512, 2000, 12)):
        outs = []
        for fuse in ("1", "0"):
            env = dict(os.environ, BK_TC_FUSE=fuse)
            r = subprocess.run(["timeout", "120", sys.executable, __file__, *map(str, shape)], env=env, capture_output=True, text=True)
            outs.append(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else f"FAILED rc={r.returncode} {r.stderr[-300:]}")
        same = outs[0] == outs[1] and not outs[0].startswith("FAILED")
        ok &= same
        print(shape, "IDENTICAL" if same else "DIFFERENT", outs[0] if same else outs, flush=True)
    sys.exit(0 if ok else 1)
