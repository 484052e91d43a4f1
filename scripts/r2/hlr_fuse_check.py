"""c3 multi-step launch vs one launch pair per leapfrog step: the draws must be bit-identical.
Runs each configuration in a subprocess with BK_HLR_FUSE=1 / 0 and compares checksums + timings."""
import json, os, subprocess, sys

HERE = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import bayes_kit_b200 as bk
from oracle.models import HierLogReg
N, Dx, C, L, n = [int(v) for v in sys.argv[1:6]]
X, y = HierLogReg.c3_data(N, Dx, seed=0)
model = bk.HierLogReg(X, y)
th0 = np.random.default_rng(1).normal(size=(C, Dx + 2)) * 0.1
s = bk.HMCDiag(model, 0.01, L, init=th0, seed=0)
d, lp = s.sample_n(n)
torch.cuda.synchronize()
h = hashlib.sha256(d.cpu().numpy().tobytes() + lp.cpu().numpy().tobytes()).hexdigest()[:16]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.sample_n(5); e1.record(); torch.cuda.synchronize()
print(json.dumps({"hash": h, "ms_per_draw": e0.elapsed_time(e1) / 5, "accept": float(s.last_accept.float().mean()),
                  "finite": bool(torch.isfinite(d).all())}))
''' % HERE

ok = True
for cfg in ((100000, 100, 1024, 10, 3), (5000, 37, 300, 4, 3), (1000, 100, 128, 7, 2), (130, 5, 50, 3, 2), (20000, 64, 2048, 5, 2)):
    res = {}
    for fuse in ("1", "0"):
        env = dict(os.environ, BK_HLR_FUSE=fuse)
        try:
            out = subprocess.run([sys.executable, "-c", CHILD] + [str(v) for v in cfg], env=env, capture_output=True,
                                 text=True, timeout=240)
            res[fuse] = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {"error": out.stderr[-400:]}
        except subprocess.TimeoutExpired:
            res[fuse] = {"error": "timeout"}
    same = "hash" in res["1"] and "hash" in res["0"] and res["1"]["hash"] == res["0"]["hash"] and res["1"].get("finite")
    ok = ok and same
    print(json.dumps({"N": cfg[0], "Dx": cfg[1], "C": cfg[2], "L": cfg[3], "identical": same, "fused": res["1"], "per_step": res["0"]}), flush=True)
print("FUSE_CHECK", "OK" if ok else "FAILED")
