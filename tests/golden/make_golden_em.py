"""Generate tests/golden/golden_em_v1.json from the CPU oracle (oracle/oracle_em.cpp).

Like golden_v1.json these vectors pin the ORACLE (and give the GPU tests a fixture that needs no
recomputation); they are NOT outputs of the Julia reference, whose SimpleEM draws from Julia's task-local
RNG.  The increments are stored with the case (hex of the IEEE bits), so regenerating the outputs involves
IEEE arithmetic only (no libm).

    python tests/golden/make_golden_em.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402

U0P = {"gbm": ([1.0], [0.1, 0.2]), "linadd1": ([0.5], [2.0, 1.0]), "linadd2": ([0.1, 0.2], [2.0, 1.0]),
       "ou": ([0.3], [1.5, 1.0, 0.4]), "nondiag2x4": ([1.0, 1.0], [1.01])}


def hexs(a):
    a = np.ascontiguousarray(a)
    return [x.tobytes().hex() for x in a.ravel()]


def unhex(h, dtype, shape):
    return np.frombuffer(bytes.fromhex("".join(h)), dtype=dtype).reshape(shape).copy()


def main():
    cases = []
    for dtype in ("float64", "float32"):
        T = np.dtype(dtype)
        for system, (u0, p) in U0P.items():
            N, NP, M, _ = O.em_dims(system)
            n, steps, t0, dt = 3, 8, 0.0, 0.125
            u0a = (np.array(u0)[:, None] * np.array([1.0, 1.25, 0.75])[None, :]).astype(T)
            pa = (np.array(p)[:, None] * np.array([1.0, 0.5, 1.5])[None, :]).astype(T)
            z = O.em_normals(T, seed=0x5eed0000 + len(cases), traj_offset=7, n_traj=n, n_steps=steps, M=M)
            out = O.em_solve(system, u0a, pa, t0, dt, steps, z)
            cases.append(dict(system=system, dtype=dtype, n=n, n_steps=steps, t0=t0, dt=dt, u0=hexs(u0a), p=hexs(pa),
                              noise=hexs(z), out=hexs(out)))
    with open(os.path.join(HERE, "golden_em_v1.json"), "w") as f:
        json.dump(dict(note="oracle_em.cpp outputs; layout out[n][n_steps+1][N], noise[n_steps][M][n], u0/p SoA",
                       cases=cases), f, indent=0)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
