"""Profiling driver: a few HMC draws of the c2 workload (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayes_kit_b200 as bk
from oracle.models import DensePrecGauss

L = int(os.environ.get("PROF_L", "4"))
C = int(os.environ.get("PROF_C", "65536"))
n = int(os.environ.get("PROF_N", "3"))
model = bk.DensePrecGauss(DensePrecGauss.c2_precision(1000, 0), dtype=torch.float32)
s = bk.HMCDiag(model, 0.1, L, chains=C, seed=0)
s.sample_n(n)
torch.cuda.synchronize()
print("done", float(s.last_accept.float().mean()))
