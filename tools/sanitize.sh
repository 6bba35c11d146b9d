#!/bin/bash
# compute-sanitizer passes over the kernels with non-trivial synchronisation / addressing
# (shared-memory staged writer, per-lane work queue).  Small sizes: the tools slow kernels ~100x.
set -u
cd "$(dirname "$0")/.."
cat > /tmp/sde_sanitize.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import simplediffeq_b200 as S, common as C
n = 1000 + 13
u0, p = C.random_problem("lorenz", n, np.float64, 1)
u0s, ps = np.ascontiguousarray(u0.T), np.ascontiguousarray(p.T)
sa = np.linspace(0, 1, 41)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.01, saveat=sa, save_mode=1, layout=0)   # staged writer
S.solve_arrays(S.systems.lorenz, S.GPUSimpleVern7(), u0s, ps, (0.0, 1.0), dt=0.05, save_mode=2, layout=0)             # staged, every step
# staged writer: many whole-line flushes per row, more save points per step than the weight ring holds, rows of every
# alignment (state sizes 1 and 2, Float32), ragged last warp
dense = S.jl_range(0.0, 0.0025, 1.0)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.0625, saveat=dense, save_mode=1, layout=0)
for name in ("vanderpol", "scalargrowth"):
    for dt_ in (np.float64, np.float32):
        a0, b0 = C.random_problem(name, 333, dt_, 2)
        S.solve_arrays(getattr(S.systems, name), S.GPUSimpleTsit5(), np.ascontiguousarray(a0.T), np.ascontiguousarray(b0.T), (0.0, 1.0),
                       dt=0.0625, saveat=S.jl_range(dt_(0.0), dt_(0.0025), dt_(1.0), dt_), save_mode=1, layout=0)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleRK4(), u0s, ps, (0.0, 1.0), dt=0.002, save_mode=2, layout=0)               # 501 slots per row
# opt-in fast kernels (SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES: stage coefficients through the kernel parameters), every save mode
S.solve_arrays(S.systems.lorenz, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.01, compat=24)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.01, saveat=sa, save_mode=1, layout=0, compat=24)
S.solve_arrays(S.systems.robertson, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.01, save_mode=2, layout=1, compat=16)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-7, reltol=1e-7)           # work queue
S.solve_arrays(S.systems.lorenz, S.GPUSimpleAVern9(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-9, reltol=1e-9, saveat=sa, save_mode=1, layout=1)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-7, reltol=1e-7, save_mode=2, out_capacity=64)
# SimpleEM: Philox stream, provided noise, non-diagonal; then the double-buffered pieces of the host path
import os
z = np.random.default_rng(0).standard_normal((16, 4, n))
S.solve_em_arrays(S.sde_systems.gbm, np.ones((1, n)), np.tile([[0.1], [0.2]], (1, n)), 0.0, 1 / 16, 16, seed=3)
S.solve_em_arrays(S.sde_systems.nondiag2x4, np.ones((2, n)), np.full((1, n), 1.01), 0.0, 1 / 16, 16, noise=z, layout=1)
S.em_noise(np.float32, 5, n, 9, 3)
# SimpleEM, every state, trajectory-major rows through the staged series writer: 1- and 2-component rows, both dtypes, ragged last warp
S.solve_em_arrays(S.sde_systems.gbm, np.ones((1, n)), np.tile([[0.1], [0.2]], (1, n)), 0.0, 1 / 256, 200, seed=3, layout=0)
S.solve_em_arrays(S.sde_systems.gbm, np.ones((1, n), dtype=np.float32), np.tile([[0.1], [0.2]], (1, n)).astype(np.float32), 0.0, 1 / 256, 200, seed=3, layout=0)
S.solve_em_arrays(S.sde_systems.nondiag2x4, np.ones((2, n)), np.full((1, n), 1.01), 0.0, 1 / 16, 16, noise=z, layout=0)
os.environ["SDE_TUNE_PIECE"] = "96"
S.solve_arrays(S.systems.lorenz, S.GPUSimpleTsit5(), u0s, ps, (0.0, 1.0), dt=0.01, saveat=sa, save_mode=1, layout=0)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-7, reltol=1e-7, saveat=sa, save_mode=1, layout=1)
S.solve_em_arrays(S.sde_systems.gbm, np.ones((1, n)), np.tile([[0.1], [0.2]], (1, n)), 0.0, 1 / 16, 16, seed=3)
print("sanitize workload done")
PY
for tool in memcheck initcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/sde_sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done|Error|error" | head -8
done
