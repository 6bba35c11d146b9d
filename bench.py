#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 ensemble ODE integrator.

Metric (BASELINE.json): trajectory-steps/s, device-timed, plus the FP64-FMA roofline fraction.
Workload at every N (weak scaling, per GPU): BASELINE.json configs[1]
    Lorenz, 10 M trajectories, GPUSimpleTsit5, fixed dt = 0.001 on tspan (0,10) -> 10 000 steps,
    FP64, endpoint only; u0 = (1,0,0), p_i = (10, rho_i, 8/3), rho_i = 21*i/(N_total-1)
    (SURVEY.md section 8d).  Trajectories shard by contiguous index range over the ranks, no
    collective on the data path.
A bench "step" = one pass of the hot path over the whole batch (= n_traj * 10 000 trajectory-steps).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [...]                         # CPU reference arm (oracle port)
    torchrun ... bench.py --gpus N ...                             # N > 1: one rank per GPU

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, CUDA-event timed, max over
ranks.  `e2e`: the same work through the C-ABI call `sde_solve` with pinned HOST buffers (H2D of
u0/p and D2H of the final states inside the timed region).  `roofline`: FP64 FMA pipe (this path is
not HBM- or tensor-bound); `roofline_hbm_config5`: the HBM-bound saveat config (configs[4]) measured
in the same run.  `cpu_baseline`: the CPU oracle (C++ restatement of the
reference; Julia is not installable here) on a bounded sample, rank 0, N = 1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_TRAJ_PER_GPU = 10_000_000
TSPAN = (0.0, 10.0)
DT = 1e-3
N_STEPS = 10_000
FLOP_PER_STEP = 190        # 63 FMA (x2) + 40 MUL + 24 ADD, dt*a21 included as the reference computes it
FP64_INSTR_PER_STEP = 127
BYTES_PER_TRAJ = 72        # 48 B in (u0, p) + 24 B out


def lorenz_inputs_np(lo, hi, n_total, dtype=np.float64):
    n = hi - lo
    u0 = np.zeros((3, n), dtype=dtype)
    u0[0] = 1
    p = np.empty((3, n), dtype=dtype)
    p[0] = 10
    p[1] = (21.0 * np.arange(lo, hi, dtype=np.float64)) / float(max(n_total - 1, 1))
    p[2] = 8.0 / 3.0
    return u0, p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        pw = []
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1])); pw.append(float(s[2]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(n_traj):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the headline kernel, from the
    committed `ncu --set full` capture -- only if that capture was taken at this launch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_bench_traffic.json")))
        if int(t["n_traj"]) == int(n_traj):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def cpu_oracle_rate(n_sample, n_threads, seconds_target=None):
    """trajectory-steps/s of the CPU oracle (C++ restatement of GPUSimpleTsit5) on a sample."""
    import oracle_lib
    from simplediffeq_b200 import jl_range
    tg = jl_range(TSPAN[0], DT, TSPAN[1])
    u0, p = lorenz_inputs_np(0, n_sample, n_sample)
    t0 = time.perf_counter()
    r = oracle_lib.solve("lorenz", "Tsit5", u0.T, p.T, TSPAN[0], TSPAN[1], DT, tgrid=tg, n_threads=n_threads)
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(r.u))
    return n_sample * N_STEPS / dt, dt


def config5_hbm_roofline(S, torch, dev, peaks, n=None, reps=5):
    """BASELINE.json configs[4] (the saveat-heavy, HBM-bound config) on this device:
    Lorenz, GPUSimpleTsit5, saveat = 0:0.01:10 (1001 points), dt = 0.1 (100 steps, 10 save points per step; the
    config does not fix dt), SoA series layout, device-resident output.  Algorithmic bytes
    per trajectory = 48 in + 1001 * 24 out = 24 072 B; achieved = bytes / kernel time
    (CUDA events on the launching stream; every launch rewrites the whole output, 96 GB >> L2)."""
    from simplediffeq_b200 import _lib, jl_range
    if n is None:      # the config's own 4 M trajectories (96.3 GB of output) when the device has the room, else 1 M
        free, _ = torch.cuda.mem_get_info(dev)
        n = 4_000_000 if free > 110e9 else 1_000_000
    u0_h, p_h = lorenz_inputs_np(0, n, n)
    d_u0, d_p = torch.from_numpy(u0_h).to(dev), torch.from_numpy(p_h).to(dev)
    saveat = jl_range(0.0, 0.01, 10.0)
    out = torch.empty((len(saveat), 3, n), dtype=torch.float64, device=dev)
    alg, sysm = S.GPUSimpleTsit5(), S.systems.lorenz
    stream = torch.cuda.current_stream(dev)

    def launch():
        S.solve_device(sysm, alg, d_u0, d_p, TSPAN, dt=0.1, saveat=saveat, save_mode=_lib.SAVE_SAVEAT,
                       layout=_lib.LAYOUT_SOA, out=out, stats=False, sync=False)
    def timed(k):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        evs[0].record(stream)
        for i in range(k):
            launch()
            evs[i + 1].record(stream)
        torch.cuda.synchronize(dev)
        return float(np.mean([evs[i].elapsed_time(evs[i + 1]) for i in range(k)]))

    # burst figure (the kernel timed alone, like MEASURED_PEAKS.json's copy bandwidth): 3 warm-up launches, then `reps`
    for _ in range(3):
        launch()
    torch.cuda.synchronize(dev)
    ms = timed(reps)
    # sustained figure: this kernel keeps the FP64 pipe ~50 % and HBM ~90 % busy at once and runs into the board's
    # power cap when launched continuously (sw_power_cap, SM clock below max): ~0.5 s of launches, then timed again
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.5:
        launch()
        torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ms_sustained = timed(2 * reps)
    clocks = sampler.summary()
    nbytes = n * (48 + len(saveat) * 24)
    gbs = nbytes / (ms * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs") or 6555.2
    finite = bool(torch.isfinite(out[-1]).all().item())
    del out
    torch.cuda.empty_cache()
    traffic = None      # dram read + write bytes of one launch from the committed ncu capture, if taken at this size
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_config5_traffic.json")))
        if int(tj["n_traj"]) == n:
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
    except Exception:
        pass
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
            "traffic_unit": "bytes per launch (ncu dram read+write; algorithmic = %d)" % nbytes,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peaks.get("hbm_gbs") else "fallback 6555.2 GB/s",
            "kernel": "sde::fixed_kernel<Lorenz,double,Tsit5Method,saveat,SoA>", "kernel_ms": ms,
            "workload": "BASELINE.json configs[4] at %d trajectories: saveat=0:0.01:10, dt=0.1, SoA output (%.1f GB per launch)" % (n, nbytes / 1e9),
            "bytes_per_trajectory": 48 + len(saveat) * 24, "launches_timed": reps,
            "sustained": {"gbs": nbytes / (ms_sustained * 1e-3) / 1e9, "kernel_ms": ms_sustained, "launches_timed": 2 * reps,
                          "after_s_of_continuous_launches": 0.5, "clocks": clocks},
            "output_finite": finite}


def config0_atsit5(S, oracle_lib, cores, dev):
    """BASELINE.json configs[0] next to the headline: Lorenz 10 k rho-sweep, GPUSimpleATsit5,
    abstol = reltol = 1e-8, tspan (0,10) -- the reference's CPU-runnable case ("GPUSimpleATsit5 under
    EnsembleThreads"): the oracle on all host threads vs the GPU through the host-buffer C-ABI call,
    in accepted trajectory-steps/s, with the step-count parity of the two."""
    n = 10_000
    u0, p = lorenz_inputs_np(0, n, n)
    dt0 = float(np.float32(0.1))
    t0 = time.perf_counter()
    o = oracle_lib.solve("lorenz", "ATsit5", u0.T, p.T, 0.0, 10.0, dt0, abstol=1e-8, reltol=1e-8, n_threads=cores)
    cpu_s = time.perf_counter() - t0
    alg = S.GPUSimpleATsit5()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        g = S.solve_arrays(S.systems.lorenz, alg, u0, p, (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8, devices=[dev.index or 0])
        best = min(best, time.perf_counter() - t0)
    acc = int(o.naccept.sum())
    return {"workload": "Lorenz 10k rho-sweep, GPUSimpleATsit5 tol 1e-8, tspan (0,10), endpoint only",
            "accepted_steps": acc, "cpu_steps_per_s": acc / cpu_s, "cpu_cores": cores, "cpu_kind": "port",
            "gpu_e2e_steps_per_s": int(g["naccept"].sum()) / best, "gpu_e2e_ms": best * 1e3,
            "identical_step_counts_frac": float(np.mean(g["naccept"] == o.naccept))}


def literal_controller_parity(S, oracle_lib, cores, dev):
    """Outside every timed region: the literal controller (SDE_COMPAT_STRICT_CONTROLLER, whose pow is the oracle's libm
    pow operation for operation) against the oracle on samples of BASELINE configs[0] (ATsit5, 1e-8) and configs[3]
    (AVern9, 1e-12 -- the configuration whose step sequence hangs on the last bit of that pow): share of trajectories
    with identical accepted AND rejected counts, share with bit-identical final states."""
    from simplediffeq_b200 import _lib
    dt0 = float(np.float32(0.1))
    out = {}
    for name, oalg, alg, n, tol in (("config0_atsit5_1e-8", "ATsit5", S.GPUSimpleATsit5(), 4096, 1e-8),
                                    ("config3_avern9_1e-12", "AVern9", S.GPUSimpleAVern9(), 4096, 1e-12)):
        u0, p = lorenz_inputs_np(0, n, n)
        o = oracle_lib.solve("lorenz", oalg, u0.T, p.T, 0.0, 10.0, dt0, abstol=tol, reltol=tol, n_threads=cores)
        g = S.solve_arrays(S.systems.lorenz, alg, u0, p, (0.0, 10.0), dt=dt0, abstol=tol, reltol=tol,
                           compat=_lib.COMPAT_STRICT_CONTROLLER, devices=[dev.index or 0])
        gu, ou = np.ascontiguousarray(g["u"].T), np.ascontiguousarray(o.u[:, 0, :])
        out[name] = {"trajectories": n,
                     "identical_step_counts_frac": float(np.mean((g["naccept"] == o.naccept) & (g["nreject"] == o.nreject))),
                     "bit_identical_final_state_frac": float(np.mean(np.all(gu.view(np.uint64) == ou.view(np.uint64), axis=1)))}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia unavailable) on all
    host threads, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    oracle_lib.build()
    cores = oracle_lib.hardware_threads()
    n_sample = 1000 * cores
    for _ in range(args.warmup):
        cpu_oracle_rate(max(n_sample // 8, cores), cores)
    tot = 0.0
    for _ in range(args.steps):
        _, dt = cpu_oracle_rate(n_sample, cores)
        tot += dt
    value = n_sample * N_STEPS * args.steps / tot
    line = {
        "impl": "reference", "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Lorenz rho-sweep, GPUSimpleTsit5 fixed dt=0.001 tspan (0,10), endpoint only (BASELINE.json configs[1])",
                   "sample": "%d trajectories x %d steps per bench step" % (n_sample, N_STEPS)},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d of 10M trajectories x %d steps, %d std::threads, g++ -O2 -mfma -ffp-contract=off"
                                   % (n_sample, N_STEPS, cores)},
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-traj", type=int, default=N_TRAJ_PER_GPU, help="trajectories per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import simplediffeq_b200 as S
    from simplediffeq_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line.  With NCCL_DEBUG=VERSION (set on the GPU boxes) NCCL printf()s its
        # "NCCL version ..." banner to stdout when the communicator is created (NCCL_DEBUG_FILE does not catch it):
        # file descriptor 1 points at stderr while the process group and its first collective are set up.
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from simplediffeq_b200.sharding import shard_bounds, endpoint_stats, gather_endpoint_stats, reduce_max
    n = args.n_traj
    n_total = n * world
    lo, hi = shard_bounds(n_total, world, rank)  # contiguous index range of this rank
    u0_h, p_h = lorenz_inputs_np(lo, hi, n_total)
    d_u0 = torch.from_numpy(u0_h).to(dev)
    d_p = torch.from_numpy(p_h).to(dev)
    d_out = torch.empty_like(d_u0)
    alg = S.GPUSimpleTsit5()
    sysm = S.systems.lorenz
    stream = torch.cuda.current_stream(dev)

    def step_device():
        S.solve_device(sysm, alg, d_u0, d_p, TSPAN, dt=DT, out=d_out, stats=False, sync=False)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- roofline denominator measured on this device, now (FP64 FMA pipe)
    peak_meas, _ = _lib.probe_fma_peak(_lib.SDE_F64)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for k in range(args.steps):
        step_device()
        evs[k + 1].record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.summary()
    ms_total = evs[0].elapsed_time(evs[-1])
    ms_kernel = float(np.mean([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]))
    ms_total = reduce_max(ms_total, dist, dev)             # max over ranks
    value = n_total * N_STEPS * args.steps / (ms_total * 1e-3)

    # ---- parity spot check of what was just computed (first 256 trajectories of rank 0)
    # ---- e2e: C-ABI call with pinned host buffers, copies inside the timed region
    u0_pin = torch.from_numpy(u0_h).pin_memory()
    p_pin = torch.from_numpy(p_h).pin_memory()
    out_pin = torch.empty((3, n), dtype=torch.float64).pin_memory()
    nacc = np.zeros(n, dtype=np.int32)
    import ctypes
    keep = []
    o = S.api.make_options(alg, np.dtype(np.float64), n, TSPAN, DT, 1e-6, 1e-3, None, _lib.SAVE_ENDPOINT,
                           _lib.LAYOUT_TRAJ_MAJOR, 0, 0, keep)
    devs = (ctypes.c_int * 1)(local)

    def step_e2e():
        rc = _lib.lib().sde_solve(sysm._handle, ctypes.byref(o), u0_pin.data_ptr(), p_pin.data_ptr(),
                                  out_pin.data_ptr(), None, None, None, None, devs, 1)
        _lib.check(rc)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_s = reduce_max(e2e_s, dist, dev)
    e2e_value = n_total * N_STEPS * e2e_steps / e2e_s
    # the e2e result must equal the device-resident result bit for bit
    assert torch.equal(out_pin, d_out.cpu()), "e2e and device-resident results differ"

    # optional final gather of endpoint statistics (the only collective; not on the data path)
    stats = gather_endpoint_stats(endpoint_stats(d_out), dist)

    if rank == 0:
        peaks = measured_peaks()
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        peak_nominal = 148 * 64 * 2 * sm_max * 1e6 / 1e12
        steps_per_s_gpu = n * N_STEPS / (ms_kernel * 1e-3)
        achieved = steps_per_s_gpu * FLOP_PER_STEP / 1e12
        line = {
            "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: Lorenz %d trajectories/GPU, GPUSimpleTsit5 fixed dt=0.001, tspan (0,10) = 10000 steps, FP64, endpoint only" % n,
                       "trajectories_total": n_total, "sharding": "contiguous index ranges, no data-path collective",
                       "l2": "inputs per pass (%d MB) exceed the 126 MB L2; the kernel is FP64-issue bound, 0.0072 B/step" % (n * 48 // 1000000)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": int(n * 48 * world),
                    "d2h_bytes_per_step": int(n * 24 * world), "api": "sde_solve (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "endpoint_stats": {"mean": [float(x) for x in stats["mean"]], "min": [float(x) for x in stats["min"]],
                               "max": [float(x) for x in stats["max"]], "gathered_over_ranks": world},
            "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
                         "frac": achieved / peak_nominal, "traffic": ncu_traffic(n),
                         "traffic_unit": "bytes per launch (ncu dram read+write; algorithmic = %d)" % (n * BYTES_PER_TRAJ),
                         "peak_source": "derived 148 SM x 64 FMA/clk x sm_max_mhz of MEASURED_PEAKS.json (no FP64 figure there); tensor cores n/a",
                         "peak_measured_dfma": peak_meas, "frac_of_measured_dfma": achieved / peak_meas if peak_meas else None,
                         "fp64_pipe_util": steps_per_s_gpu * FP64_INSTR_PER_STEP / (148 * 64 * sm_max * 1e6),
                         "kernel": "sde::fixed_kernel<Lorenz,double,Tsit5Method,endpoint>", "kernel_ms": ms_kernel,
                         "flop_per_step": FLOP_PER_STEP, "fp64_instr_per_step": FP64_INSTR_PER_STEP,
                         "hbm_gbs": n * BYTES_PER_TRAJ / (ms_kernel * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle_lib
            oracle_lib.build()
            cores = oracle_lib.hardware_threads()
            n_s = 4000 * cores          # ~10-15 s of wall time on the box's host cores
            rate, secs = cpu_oracle_rate(n_s, cores)
            # parity: the oracle's sample is the first n_s trajectories of a different sweep; check a slice directly
            u0c, pc = lorenz_inputs_np(0, 512, n_total)
            from simplediffeq_b200 import jl_range
            r = oracle_lib.solve("lorenz", "Tsit5", u0c.T, pc.T, TSPAN[0], TSPAN[1], DT, tgrid=jl_range(TSPAN[0], DT, TSPAN[1]), n_threads=cores)
            same = bool(np.array_equal(r.u[:, 0, :].T, d_out[:, :512].cpu().numpy()))
            line["cpu_baseline"] = {"value": rate, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d of 10M trajectories x 10000 steps in %.1f s, %d std::threads (C++ restatement of GPUSimpleTsit5; Julia unavailable)" % (n_s, secs, cores)}
            line["parity_spot_check"] = {"trajectories": 512, "bit_identical_to_oracle": same}
            line["config0_atsit5"] = config0_atsit5(S, oracle_lib, cores, dev)
            try:        # parity evidence only, after the timed region: reported, never fatal
                line["literal_controller_parity"] = literal_controller_parity(S, oracle_lib, cores, dev)
            except Exception as e:
                line["literal_controller_parity"] = {"error": str(e)[:200]}
            try:        # second roofline of BASELINE.json's metric ("% FP64 FMA / HBM roofline"): the saveat-heavy config
                line["roofline_hbm_config5"] = config5_hbm_roofline(S, torch, dev, peaks)
            except Exception as e:    # e.g. not enough free memory next to another tenant: reported, never fatal
                line["roofline_hbm_config5"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
