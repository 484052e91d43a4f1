"""One launch of the c1 workload (for ncu): python scripts/r2/one_call_c1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bayes_kit_b200 as bk
s = bk.HMCDiag(bk.IsoGauss(100), 0.1, 10, chains=1048576, seed=0)
s.sample_n(10)
torch.cuda.synchronize()
