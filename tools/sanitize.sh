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
S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-7, reltol=1e-7)           # work queue
S.solve_arrays(S.systems.lorenz, S.GPUSimpleAVern9(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-9, reltol=1e-9, saveat=sa, save_mode=1, layout=1)
S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), u0s, ps, (0.0, 1.0), dt=0.1, abstol=1e-7, reltol=1e-7, save_mode=2, out_capacity=64)
print("sanitize workload done")
PY
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/sde_sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done|Error|error" | head -8
done
