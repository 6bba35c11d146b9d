"""BASELINE config 1 at its own size (10 k Lorenz trajectories, GPUSimpleATsit5 tol 1e-8): what bounds the 0.6 ms kernel?
Device time of the persistent adaptive kernel for sub-ensembles made of the LONGEST trajectories of the sweep (largest rho):
1 trajectory, one warp, one CTA per SM, ... the whole sweep.  If one trajectory alone already takes most of the 10 k
ensemble's time, the ensemble is bound by the serial dependency chain of its longest member (attempts x cycles per
attempt), which no launch geometry can shorten.
    python tools/small_latency.py > gpurun_out/small_latency.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import simplediffeq_b200 as S  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n_all = 10_000
    rho = 21.0 * np.arange(n_all, dtype=np.float64) / (n_all - 1)
    dt0 = float(np.float32(0.1))
    alg = S.GPUSimpleATsit5()

    def run(idx, compat=0, reps=20):
        m = len(idx)
        u0 = torch.zeros(3, m, dtype=torch.float64, device=dev); u0[0] = 1
        p = torch.empty(3, m, dtype=torch.float64, device=dev); p[0] = 10; p[2] = 8.0 / 3.0
        p[1] = torch.from_numpy(rho[idx]).to(dev)
        out = None
        ts = []
        for r in range(reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = S.solve_device(S.systems.lorenz, alg, u0, p, (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8, compat=compat, sync=False)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        att = (out["naccept"] + out["nreject"])
        return float(np.median(ts[3:])), int(att.max().item()), float(att.double().mean().item())

    full = np.arange(n_all)
    order = None
    ms, amax, amean = run(full)
    print("whole sweep (10 000 trajectories): %.3f ms; attempts per trajectory: mean %.1f, max %d" % (ms, amean, amax))
    # the longest trajectories first
    u0 = torch.zeros(3, n_all, dtype=torch.float64, device=dev); u0[0] = 1
    p = torch.empty(3, n_all, dtype=torch.float64, device=dev); p[0] = 10; p[2] = 8.0 / 3.0
    p[1] = torch.from_numpy(rho).to(dev)
    o = S.solve_device(S.systems.lorenz, alg, u0, p, (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8)
    att = (o["naccept"] + o["nreject"]).cpu().numpy()
    order = np.argsort(-att)
    clk = 1.965e9
    for m in (1, 32, 128, 148 * 32, 148 * 128, n_all):
        idx = np.sort(order[:m])
        ms, amax, amean = run(idx)
        print("%6d longest trajectories: %.3f ms  (max attempts %d -> %.0f cycles per attempt of the longest at 1965 MHz; mean attempts %.0f)"
              % (m, ms, amax, ms * 1e-3 * clk / amax, amean))
    ms, amax, _ = run(np.sort(order[:1]), compat=2)
    print("     1 longest trajectory, literal controller: %.3f ms (%.0f cycles per attempt)" % (ms, ms * 1e-3 * clk / amax))
    # an empty launch for scale
    ms, _, _ = run(np.sort(order[-1:]))
    print("     1 shortest trajectory: %.3f ms" % ms)


if __name__ == "__main__":
    main()
