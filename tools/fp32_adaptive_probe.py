import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import simplediffeq_b200 as S
from quick_bench import probe
dt0 = float(np.float32(0.1))
for compat, label in ((0, "default(literal)"), (4, "log2 forced")):
    for tol in (1e-4, 1e-6):
        print(label, tol, end=" ")
        probe("lorenz", S.GPUSimpleATsit5(), 1 << 20, (0.0, 10.0), dt0, dtype=torch.float32, abstol=tol, reltol=tol, compat=compat)
    print(label, end=" ")
    probe("vanderpol", S.GPUSimpleATsit5(), 1 << 20, (0.0, 20.0), dt0, dtype=torch.float32, abstol=1e-5, reltol=1e-5, compat=compat)
    print(label, end=" ")
    probe("lorenz", S.GPUSimpleAVern7(), 1 << 20, (0.0, 10.0), dt0, dtype=torch.float32, abstol=1e-6, reltol=1e-6, compat=compat)
