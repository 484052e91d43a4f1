"""One sample_n(n) call of the c2 workload (for ncu launch lists): python scripts/r2/one_call_c2.py [n] [L]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
from oracle.models import DensePrecGauss
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model = bk.DensePrecGauss(DensePrecGauss.c2_precision(1000, 0), dtype=torch.float32)
s = bk.HMCDiag(model, 0.1, L, chains=65536, seed=0)
s.sample_n(n)
torch.cuda.synchronize()
