"""Development aid: run one adaptive configuration (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import simplediffeq_b200 as S
import bench_configs as B
which = sys.argv[1]
COMPAT = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # 2 = literal controller
if which == "atsit5":
    u0, p = B.lorenz(1 << 19)
    for _ in range(2):
        S.solve_device(S.systems.lorenz, S.GPUSimpleATsit5(), u0, p, (0.0, 10.0), dt=B.DT0, abstol=1e-8, reltol=1e-8, compat=COMPAT)
elif which == "vdp":
    u0, p = B.vdp(1 << 19, True)
    for _ in range(2):
        S.solve_device(S.systems.vanderpol, S.GPUSimpleATsit5(), u0, p, (0.0, 20.0), dt=B.DT0, abstol=1e-6, reltol=1e-6, compat=COMPAT)
elif which == "avern9":
    u0, p = B.lorenz(1 << 18)
    for _ in range(2):
        S.solve_device(S.systems.lorenz, S.GPUSimpleAVern9(), u0, p, (0.0, 10.0), dt=B.DT0, abstol=1e-12, reltol=1e-12, compat=COMPAT)
