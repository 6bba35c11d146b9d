"""Development aid: one SimpleEM launch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simplediffeq_b200 as S
dev = torch.device("cuda:0")
dtype = torch.float32 if "f32" in sys.argv else torch.float64
n, steps = 1 << 21, 500
u0 = torch.ones((1, n), dtype=dtype, device=dev)
p = torch.empty((2, n), dtype=dtype, device=dev); p[0] = 0.1; p[1] = 0.2
for _ in range(2):
    S.solve_em_device(S.sde_systems.gbm, u0, p, 0.0, 1e-3, steps, seed=1)
