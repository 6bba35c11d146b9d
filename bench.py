#!/usr/bin/env python3
"""bench.py -- benchmark of the B200 ensemble ODE integrator (libsimplediffeq_cuda).

Metric (BASELINE.json): trajectory-steps/s (accepted steps), device-timed, plus the FP64-FMA / HBM
roofline fraction.  A bench "step" = one pass of the hot path over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--scaling weak|strong]
    python bench.py --impl reference [...]                # CPU reference arm (oracle port, all host threads)
    torchrun ... bench.py --gpus N ...                    # N > 1: one rank per GPU, trajectories sharded

--config selects the contract line's workload (default 2 = BASELINE.json configs[1],
the configuration the headline metric is quoted on):
    1     Lorenz rho-sweep, 10 k trajectories, GPUSimpleATsit5 tol 1e-8       (configs[0], the CPU-runnable case)
    1b    the same sweep at 2^20 trajectories (a GPU-filling ensemble)
    2     Lorenz 10 M, GPUSimpleTsit5 fixed dt = 1e-3, 10 000 steps, FP64     (configs[1])   [default]
    2f32  the same in Float32
    2fast the same in FP64 with SDE_COMPAT_FAST_RHS (contracted right-hand side; not bit-identical to the reference)
    2faster  ... and SDE_COMPAT_FAST_STAGES (step size folded into the stage coefficients)
    3     Van der Pol mu-sweep 2^20, GPUSimpleATsit5 tol 1e-6, sorted         (configs[2])
    3s    the same, shuffled (i -> i * 2654435761 mod n)
    4     Lorenz 1 M, GPUSimpleAVern9 tol 1e-12 (default options = literal controller)   (configs[3])
    4l    the same with the log2-domain controller forced (not step-count compliant, for comparison)
    5     Lorenz 4 M, GPUSimpleTsit5 + saveat = 0:0.01:10, dt = 0.1, SoA series output   (configs[4], HBM bound)
    5tm   the same, trajectory-major output
    5f    the same at dt = 0.01 (one save point per step: FP64 bound), SoA
The default run (config 2) also times every other config on the same device(s) -- device-resident, CUDA events,
max over ranks, each with its own roofline -- and reports them under "configs"; at N > 1 every config is SHARDED
by contiguous index range (strong scaling of the config's own ensemble, and weak scaling where stated), adaptive
sorted sweeps additionally with the cost-weighted split of simplediffeq_b200.sharding, config 2 additionally in
strong scaling (10 M trajectories total), and rank 0 drives ALL devices once through the in-library sharder
`sde_solve(..., devices = [0..N-1])` (the route the Julia shim takes) while the other ranks wait on a CPU barrier.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, CUDA-event timed on the launching stream,
max over ranks.  `e2e`: the same work through the C-ABI call `sde_solve` with pinned HOST buffers (H2D of u0 / p and
D2H of the results inside the timed region).  `cpu_baseline`: the CPU oracle (C++ restatement of the reference; Julia
is not installable here) on a bounded sample, rank 0, N = 1 only.
"""
import argparse
import ctypes
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

DT0 = float(np.float32(0.1))       # the reference's default dt = 0.1f0
SM_COUNT, FP64_LANES, FP32_LANES = 148, 64, 128

# name -> workload.  instr = FP64 (FP32) pipe instructions per step / attempt of the shipped kernel,
# flop = the flops among them (FMA = 2)
CONFIGS = {
    "1": dict(baseline="configs[0]", system="lorenz", alg="GPUSimpleATsit5", oalg="ATsit5", n=10_000, tspan=(0.0, 10.0), tol=1e-8,
              desc="Lorenz rho-sweep 10k trajectories, GPUSimpleATsit5 abstol=reltol=1e-8, tspan (0,10), endpoint only", instr=218),
    "1b": dict(baseline="configs[0] sweep at 2^20", system="lorenz", alg="GPUSimpleATsit5", oalg="ATsit5", n=1 << 20, tspan=(0.0, 10.0), tol=1e-8,
               desc="Lorenz rho-sweep 2^20 trajectories, GPUSimpleATsit5 abstol=reltol=1e-8, tspan (0,10), endpoint only", instr=218),
    "2": dict(baseline="configs[1]", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=10_000_000, tspan=(0.0, 10.0), dt=1e-3,
              desc="Lorenz rho-sweep 10M trajectories, GPUSimpleTsit5 fixed dt=0.001, tspan (0,10) = 10000 steps, FP64, endpoint only",
              instr=127, flop=190),
    "2f32": dict(baseline="configs[1] in Float32", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=10_000_000, tspan=(0.0, 10.0), dt=1e-3,
                 dtype="f32", desc="Lorenz rho-sweep 10M trajectories, GPUSimpleTsit5 fixed dt=0.001, 10000 steps, FP32, endpoint only",
                 instr=127, flop=190),
    "2fast": dict(baseline="configs[1] with SDE_COMPAT_FAST_RHS", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=10_000_000,
                  tspan=(0.0, 10.0), dt=1e-3, compat=8, instr=115, flop=190,
                  desc="Lorenz rho-sweep 10M trajectories, GPUSimpleTsit5 fixed dt=0.001, FP64, endpoint only, contracted right-hand side "
                       "(SDE_COMPAT_FAST_RHS: 114 instead of 126 FP64 operations per step; NOT bit-identical to the reference, <= 1e-12 "
                       "relative on this sweep; the same 190 flops per step are counted)"),
    "2faster": dict(baseline="configs[1] with SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5",
                    n=10_000_000, tspan=(0.0, 10.0), dt=1e-3, compat=24, instr=99, flop=190,
                    desc="Lorenz rho-sweep 10M trajectories, GPUSimpleTsit5 fixed dt=0.001, FP64, endpoint only, contracted right-hand side and "
                         "the step size folded into the stage coefficients (SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES: 99 instead of 126 "
                         "FP64 instructions per step; NOT bit-identical to the reference: median 5e-15 relative on this sweep, <= 1e-12 "
                         "away from the bifurcation at rho = 13.926; the reference method's 190 flops per step are counted)"),
    "3": dict(baseline="configs[2] sorted", system="vanderpol", alg="GPUSimpleATsit5", oalg="ATsit5", n=1 << 20, tspan=(0.0, 20.0), tol=1e-6,
              desc="Van der Pol mu-sweep 2^20 trajectories (sorted), GPUSimpleATsit5 abstol=reltol=1e-6, tspan (0,20), endpoint only", instr=150),
    "3s": dict(baseline="configs[2] shuffled", system="vanderpol", alg="GPUSimpleATsit5", oalg="ATsit5", n=1 << 20, tspan=(0.0, 20.0), tol=1e-6,
               shuffled=True,
               desc="Van der Pol mu-sweep 2^20 trajectories (shuffled i -> i*2654435761 mod n), GPUSimpleATsit5 tol 1e-6, endpoint only", instr=150),
    "4": dict(baseline="configs[3]", system="lorenz", alg="GPUSimpleAVern9", oalg="AVern9", n=1_000_000, tspan=(0.0, 10.0), tol=1e-12,
              desc="Lorenz rho-sweep 1M trajectories, GPUSimpleAVern9 abstol=reltol=1e-12, endpoint only, DEFAULT options "
                   "(reltol <= 1e-11 selects the literal controller: step counts identical to the CPU oracle)", instr=620),
    "4l": dict(baseline="configs[3], log2-domain controller forced", system="lorenz", alg="GPUSimpleAVern9", oalg="AVern9", n=1_000_000,
               tspan=(0.0, 10.0), tol=1e-12, compat=4,
               desc="Lorenz rho-sweep 1M trajectories, GPUSimpleAVern9 tol 1e-12, SDE_COMPAT_LOG2_CONTROLLER (step counts differ from "
                    "the oracle on ~half of the trajectories: comparison only)", instr=529),
    "5": dict(baseline="configs[4]", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=4_000_000, tspan=(0.0, 10.0), dt=0.1,
              saveat=(0.0, 0.01, 10.0), layout="soa",
              desc="Lorenz 4M trajectories, GPUSimpleTsit5 dt=0.1 + saveat=0:0.01:10 (1001 points), SoA series output (96.3 GB)"),
    "5tm": dict(baseline="configs[4], trajectory-major", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=4_000_000, tspan=(0.0, 10.0),
                dt=0.1, saveat=(0.0, 0.01, 10.0), layout="traj_major",
                desc="Lorenz 4M trajectories, GPUSimpleTsit5 dt=0.1 + saveat=0:0.01:10, trajectory-major series output (96.3 GB)"),
    "5f": dict(baseline="configs[4] at dt=0.01", system="lorenz", alg="GPUSimpleTsit5", oalg="Tsit5", n=4_000_000, tspan=(0.0, 10.0), dt=0.01,
               saveat=(0.0, 0.01, 10.0), layout="soa", instr=151,
               desc="Lorenz 4M trajectories, GPUSimpleTsit5 dt=0.01 (1000 steps, ~1 save point per step) + saveat=0:0.01:10, SoA"),
}
EXTRA_ORDER = ["2f32", "2fast", "2faster", "1", "1b", "3", "3s", "4", "4l", "5", "5f", "5tm"]


def is_adaptive(cfg):
    return "tol" in cfg


def np_dtype(cfg):
    return np.float32 if cfg.get("dtype") == "f32" else np.float64


def n_steps_of(cfg):
    from simplediffeq_b200 import jl_range
    return len(jl_range(cfg["tspan"][0], cfg["dt"], cfg["tspan"][1])) - 1


def bytes_per_traj(cfg):
    es = 4 if cfg.get("dtype") == "f32" else 8
    n_state, n_par = (2, 1) if cfg["system"] == "vanderpol" else (3, 3)
    n_save = 1
    if "saveat" in cfg:
        a, s, b = cfg["saveat"]
        n_save = int(round((b - a) / s)) + 1
    return es * (n_state + n_par + n_state * n_save)


# ------------------------------------------------------------------------------------------------
# inputs: deterministic, the same formulas in numpy (host) and torch (device)
# ------------------------------------------------------------------------------------------------
def inputs_np_idx(cfg, idx, n_total):
    """SoA inputs of the trajectories with global indices `idx` (int64 array) of a sweep of n_total."""
    dtype = np_dtype(cfg)
    n = len(idx)
    if cfg.get("shuffled"):
        idx = (idx * 2654435761) % n_total
    den = float(max(n_total - 1, 1))
    if cfg["system"] == "lorenz":
        u0 = np.zeros((3, n), dtype=dtype)
        u0[0] = 1
        p = np.empty((3, n), dtype=dtype)
        p[0] = 10
        p[1] = ((21.0 * idx.astype(np.float64)) / den).astype(dtype)
        p[2] = dtype(8.0) / dtype(3.0) if dtype is np.float32 else 8.0 / 3.0
    else:
        u0 = np.zeros((2, n), dtype=dtype)
        u0[0] = 2
        p = (0.1 + 49.9 * (idx.astype(np.float64) / den)).astype(dtype).reshape(1, n)
    return u0, np.ascontiguousarray(p)


def inputs_np(cfg, lo, hi, n_total):
    return inputs_np_idx(cfg, np.arange(lo, hi, dtype=np.int64), n_total)


def lorenz_inputs_np(lo, hi, n_total, dtype=np.float64):
    """Config 2's inputs for the index range [lo, hi) of a sweep of n_total (also used by tests/test_sharding_gloo.py)."""
    return inputs_np(CONFIGS["2f32" if dtype is np.float32 else "2"], lo, hi, n_total)


def inputs_torch(cfg, lo, hi, n_total, dev, torch):
    """The same arrays on the device.  Built on the host and copied: torch divides by a Python scalar as a multiplication
    by its reciprocal on the GPU, which moves rho by one ulp -- and a chaotic Lorenz trajectory by much more."""
    u0, p = inputs_np(cfg, lo, hi, n_total)
    return torch.from_numpy(u0).to(dev), torch.from_numpy(p).to(dev)


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


def committed_traffic(fname, n_traj):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from a COMMITTED `ncu --set full` capture (not a live
    measurement: ncu cannot run inside the timed bench) -- only if that capture was taken at this launch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", fname)))
        if int(t["n_traj"]) == int(n_traj):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# the CPU arm: the oracle (C++ restatement of the reference's solve methods) on a sample of a config
# ------------------------------------------------------------------------------------------------
def cpu_oracle_run(cfg, n_sample, n_threads):
    """(accepted trajectory-steps, seconds, result) of the CPU oracle on the first n_sample trajectories of a sweep of
    n_sample (the same formulas as the GPU inputs)."""
    import oracle_lib
    from simplediffeq_b200 import jl_range
    u0, p = inputs_np(cfg, 0, n_sample, n_sample)
    dtype = np_dtype(cfg)
    kw = dict(dtype=dtype, n_threads=n_threads)
    if is_adaptive(cfg):
        t0 = time.perf_counter()
        r = oracle_lib.solve(cfg["system"], cfg["oalg"], u0.T, p.T, cfg["tspan"][0], cfg["tspan"][1], DT0, abstol=cfg["tol"],
                             reltol=cfg["tol"], **kw)
        secs = time.perf_counter() - t0
        return int(r.naccept.sum()), secs, r
    tg = jl_range(dtype(cfg["tspan"][0]), dtype(cfg["dt"]), dtype(cfg["tspan"][1]), dtype)
    if "saveat" in cfg:
        kw["saveat"] = jl_range(*cfg["saveat"]).astype(dtype)
    t0 = time.perf_counter()
    r = oracle_lib.solve(cfg["system"], cfg["oalg"], u0.T, p.T, cfg["tspan"][0], cfg["tspan"][1], cfg["dt"], tgrid=tg, **kw)
    secs = time.perf_counter() - t0
    assert np.all(np.isfinite(r.u))
    return n_sample * (len(tg) - 1), secs, r


def cpu_sample_size(cfg, cores):
    """A sample worth roughly 10 s of CPU work on `cores` threads (measured oracle rates per thread: ~9e6 steps/s fixed
    Tsit5, ~5e6 accepted steps/s ATsit5 at ~480 steps per trajectory, ~1.5e6 AVern9), bounded by the config's own
    ensemble and by 2 GB of series output."""
    if cfg["alg"] == "GPUSimpleTsit5":
        work = n_steps_of(cfg) + (500 if "saveat" in cfg else 0)
        n = int(1e8 * cores / work)
    elif cfg["alg"] == "GPUSimpleAVern9":
        n = 30000 * cores
    else:
        n = 100000 * cores
    return max(cores, min(n, cfg["n"], int(2e9 / bytes_per_traj(cfg))))


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia unavailable) on all host threads, one
    bounded sample of the selected config per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    oracle_lib.build()
    cfg = CONFIGS[args.config]
    cores = oracle_lib.hardware_threads()
    n_sample = min(cfg["n"], max(cores, cpu_sample_size(cfg, cores) // 10))     # ~1-2 s per bench step
    for _ in range(args.warmup):
        cpu_oracle_run(cfg, max(n_sample // 8, cores), cores)
    tot, steps_done = 0.0, 0
    for _ in range(args.steps):
        acc, dt, _ = cpu_oracle_run(cfg, n_sample, cores)
        tot += dt
        steps_done += acc
    value = steps_done / tot
    line = {
        "impl": "reference", "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": cfg.get("dtype", "f64"), "data": "synthetic",
        "config": {"workload": "BASELINE.json %s: %s" % (cfg["baseline"], cfg["desc"]),
                   "sample": "%d trajectories per bench step" % n_sample},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d trajectories per step, %d std::threads, g++ -O2 -mfma -ffp-contract=off (C++ restatement of "
                                   "the reference's solve method; Julia unavailable)" % (n_sample, cfg["n"], cores)},
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import simplediffeq_b200 as S
        from simplediffeq_b200 import _lib
        self.torch, self.S, self._lib, self.args = torch, S, _lib, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        self.cpu_group = None
        if self.world > 1:
            # stdout carries exactly one JSON line.  With NCCL_DEBUG=VERSION (set on the GPU boxes) NCCL printf()s its
            # "NCCL version ..." banner to stdout when the communicator is created (NCCL_DEBUG_FILE does not catch it):
            # file descriptor 1 points at stderr while the process group and its first collective are set up.
            import torch.distributed as dist
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize(self.dev)
                # a CPU-side group: ranks that must keep their GPU idle (while rank 0 drives all devices through the
                # in-library sharder) wait here instead of spinning in an NCCL kernel
                self.cpu_group = dist.new_group(backend="gloo")
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
            self.dist = dist
        self.peaks = measured_peaks()
        self.sm_max = self.peaks.get("sm_max_mhz", 1965.0)
        self.stream = torch.cuda.current_stream(self.dev)

    # ---- plumbing
    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, value, op="max"):
        if self.dist is None:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def pipe_peak(self, cfg):
        lanes = FP32_LANES if cfg.get("dtype") == "f32" else FP64_LANES
        return SM_COUNT * lanes * self.sm_max * 1e6          # pipe instructions / s

    # ---- one config on this rank's shard [lo, hi) of a sweep of n_total
    def make_runner(self, cfg, lo, hi, n_total):
        torch, S, _lib = self.torch, self.S, self._lib
        sysm = getattr(S.systems, cfg["system"])
        alg = getattr(S, cfg["alg"])()
        d_u0, d_p = inputs_torch(cfg, lo, hi, n_total, self.dev, torch)
        n = hi - lo
        st = {"out": None}
        if is_adaptive(cfg):
            def launch():
                st["out"] = S.solve_device(sysm, alg, d_u0, d_p, cfg["tspan"], dt=DT0, abstol=cfg["tol"], reltol=cfg["tol"],
                                           compat=cfg.get("compat", 0), sync=False)
        elif "saveat" in cfg:
            saveat = S.jl_range(*cfg["saveat"]).astype(np_dtype(cfg))
            lay = _lib.LAYOUT_SOA if cfg["layout"] == "soa" else _lib.LAYOUT_TRAJ_MAJOR
            shape = (len(saveat), 3, n) if cfg["layout"] == "soa" else (n, len(saveat), 3)
            out = torch.empty(shape, dtype=d_u0.dtype, device=self.dev)

            def launch():
                st["out"] = S.solve_device(sysm, alg, d_u0, d_p, cfg["tspan"], dt=cfg["dt"], saveat=saveat, save_mode=_lib.SAVE_SAVEAT,
                                           layout=lay, out=out, stats=False, sync=False)
        else:
            out = torch.empty_like(d_u0)

            def launch():
                st["out"] = S.solve_device(sysm, alg, d_u0, d_p, cfg["tspan"], dt=cfg["dt"], out=out, stats=False, sync=False,
                                           compat=cfg.get("compat", 0))
        return launch, st, (d_u0, d_p)

    def time_launches(self, launch, warmup, steps, sample_clocks=False):
        """`warmup` untimed launches, then `steps` launches bracketed by barrier + synchronize; per-launch CUDA events on the
        launching stream.  Returns (ms for all steps = max over ranks, mean ms per launch on this rank, clocks)."""
        torch = self.torch
        for _ in range(warmup):
            launch()
        self.barrier()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(self.local)
            sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        self.barrier()
        evs[0].record(self.stream)
        for k in range(steps):
            launch()
            evs[k + 1].record(self.stream)
        self.barrier()
        clocks = sampler.summary() if sampler else None
        ms_total = self.reduce(evs[0].elapsed_time(evs[-1]))
        ms_kernel = float(np.mean([evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]))
        return ms_total, ms_kernel, clocks

    def work_of(self, cfg, st, n_local):
        """(accepted steps, attempts) of one launch on this rank."""
        if is_adaptive(cfg):
            o = st["out"]
            acc = int(o["naccept"].sum().item())
            return acc, acc + int(o["nreject"].sum().item())
        w = n_local * n_steps_of(cfg)
        return w, w

    def roofline_of(self, cfg, steps_per_s_gpu, attempts_per_s_gpu, n_local, ms_kernel):
        """The roofline that bounds this config's kernel, from ONE GPU's rate (rank 0's shard)."""
        hbm_peak = self.peaks.get("hbm_gbs") or 6650.0
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if self.peaks.get("hbm_gbs") else "fallback 6650 GB/s"
        gbs = n_local * bytes_per_traj(cfg) / (ms_kernel * 1e-3) / 1e9
        if "saveat" in cfg and "instr" not in cfg:
            return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src,
                    "bytes_per_trajectory": bytes_per_traj(cfg),
                    "traffic": committed_traffic("r2_ncu_config5tm_traffic.json" if cfg.get("layout") == "traj_major" else "r2_ncu_config5_traffic.json", n_local),
                    "traffic_source": "committed ncu --set full capture of this launch size (profiles/r2_ncu_config5_*.summary.txt), not a live measurement"}
        pipe = self.pipe_peak(cfg)
        r = {"bound": "fp64_pipe" if cfg.get("dtype") != "f32" else "fp32_pipe",
             "achieved": attempts_per_s_gpu * cfg["instr"] / 1e12, "peak": pipe / 1e12, "unit": "T pipe-instr/s",
             "frac": attempts_per_s_gpu * cfg["instr"] / pipe, "instr_per_attempt": cfg["instr"],
             "peak_source": "148 SM x %d lanes x sm_max_mhz of MEASURED_PEAKS.json" % (FP32_LANES if cfg.get("dtype") == "f32" else FP64_LANES),
             "hbm_gbs": gbs}
        if "flop" in cfg:
            r["flop_frac_of_fma_peak"] = steps_per_s_gpu * cfg["flop"] / (2 * pipe)
        return r

    def run_config(self, name, n_total=None, bounds=None, warmup=3, steps=5, split="equal index ranges"):
        """One config sharded over the ranks; returns the result dict (identical on every rank except rank-local detail)."""
        from simplediffeq_b200.sharding import shard_bounds
        cfg = CONFIGS[name]
        n_total = cfg["n"] if n_total is None else n_total
        lo, hi = (bounds[self.rank], bounds[self.rank + 1]) if bounds is not None else shard_bounds(n_total, self.world, self.rank)
        launch, st, keep = self.make_runner(cfg, lo, hi, n_total)
        hbm_bound = "saveat" in cfg
        if hbm_bound:
            # the HBM-bound kernels run into the board's power cap when they follow seconds of FP64-saturating launches
            # (round 1: 0.835 of peak inside the bench against 0.90 alone): two seconds of idle first, clocks sampled
            # during the timed launches, so that the figure is the kernel's and the clock record says under what conditions
            self.torch.cuda.synchronize(self.dev)
            time.sleep(2.0)
        ms_total, ms_kernel, clocks = self.time_launches(launch, warmup, steps, sample_clocks=hbm_bound)
        acc, att = self.work_of(cfg, st, hi - lo)
        acc_all, att_all = self.reduce(acc, "sum"), self.reduce(att, "sum")
        res = {"workload": cfg["desc"], "baseline": cfg["baseline"], "trajectories_total": n_total, "n_gpus": self.world,
               "split": split, "shard_sizes": [int(self.reduce(hi - lo if r == self.rank else 0, "sum")) for r in range(self.world)]
               if self.world > 1 else [hi - lo],
               "ms": ms_total / steps, "launches_timed": steps,
               "trajectory_steps_per_s": acc_all * steps / (ms_total * 1e-3)}
        if is_adaptive(cfg):
            res["attempts_per_s"] = att_all * steps / (ms_total * 1e-3)
            res["accepted_steps"] = int(acc_all)
            res["reject_frac"] = 1.0 - acc_all / max(att_all, 1.0)
            res["controller"] = ("literal" if (cfg.get("compat", 0) == 2 or (cfg.get("compat", 0) == 0 and (cfg["tol"] <= 1e-11 or cfg.get("dtype") == "f32")))
                                 else "log2-domain")
            failed = int(self.reduce(int((st["out"]["retcode"] != 0).sum().item()), "sum"))
            res["failed"] = failed
        rl = self.roofline_of(cfg, acc / (ms_kernel * 1e-3), att / (ms_kernel * 1e-3), hi - lo, ms_kernel)
        rl["kernel_ms_rank0"] = ms_kernel
        res["roofline"] = rl
        if "saveat" in cfg:
            res["hbm_gbs_aggregate"] = n_total * bytes_per_traj(cfg) * steps / (ms_total * 1e-3) / 1e9
            res["clocks"] = clocks
            res["timing_note"] = "after 2 s of idle (burst conditions, like the copy benchmark behind MEASURED_PEAKS.json hbm_gbs)"
        self._last = (cfg, st, keep, lo, hi)
        return res

    def free(self):
        self._last = None
        self.torch.cuda.empty_cache()

    # ---- the host-buffer C-ABI call on this rank's shard (endpoint-only configs)
    def e2e_of(self, name, lo, hi, n_total, d_ref, steps):
        torch, S, _lib = self.torch, self.S, self._lib
        cfg = CONFIGS[name]
        if "saveat" in cfg:
            return None
        dtype = np.dtype(np_dtype(cfg))
        n = hi - lo
        u0_h, p_h = inputs_np(cfg, lo, hi, n_total)
        u0_pin, p_pin = torch.from_numpy(u0_h).pin_memory(), torch.from_numpy(p_h).pin_memory()
        out_pin = torch.empty((u0_h.shape[0], n), dtype=torch.from_numpy(u0_h).dtype).pin_memory()
        sysm, alg = getattr(S.systems, cfg["system"]), getattr(S, cfg["alg"])()
        keep = []
        adaptive = is_adaptive(cfg)
        o = S.api.make_options(alg, dtype, n, cfg["tspan"], DT0 if adaptive else cfg["dt"], cfg.get("tol", 1e-6), cfg.get("tol", 1e-3), None,
                               _lib.SAVE_ENDPOINT, _lib.LAYOUT_TRAJ_MAJOR, cfg.get("compat", 0), 0, keep)
        nacc = torch.zeros(n, dtype=torch.int32).pin_memory() if adaptive else None
        devs = (ctypes.c_int * 1)(self.local)

        def step():
            rc = _lib.lib().sde_solve(sysm._handle, ctypes.byref(o), u0_pin.data_ptr(), p_pin.data_ptr(), out_pin.data_ptr(), None,
                                      nacc.data_ptr() if adaptive else None, None, None, devs, 1)
            _lib.check(rc)

        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()                           # synchronous: returns when the results are in the host buffers
        self.barrier()
        secs = self.reduce(time.perf_counter() - t0)
        same = bool(torch.equal(out_pin, d_ref.cpu())) if d_ref is not None else None
        acc = int(nacc.sum().item()) if adaptive else n * n_steps_of(cfg)
        acc_all = self.reduce(acc, "sum")
        es = dtype.itemsize
        return {"value": acc_all * steps / secs, "unit": "trajectory-steps/s",
                "h2d_bytes_per_step": int(n_total * es * (u0_h.shape[0] + p_h.shape[0])),
                "d2h_bytes_per_step": int(n_total * (es * u0_h.shape[0] + (4 if adaptive else 0))),
                "api": "sde_solve (C ABI, pinned host buffers, one call per step per rank)",
                "calls_timed": steps, "timer": "time.perf_counter around the synchronous calls, barrier on both sides, max over ranks",
                "bit_identical_to_device_resident_result": same}

    # ---- rank 0 drives every device of the box through the in-library sharder (the Julia shim's route)
    def inlib_sharder(self, name, n_per_gpu, per_rank_out, calls=3):
        torch, S, _lib, dist = self.torch, self.S, self._lib, self.dist
        cfg = CONFIGS[name]
        n_total = n_per_gpu * self.world
        # every rank: 64-bit checksum of its device-resident result (bit patterns summed mod 2^63)
        my_sum = int(per_rank_out.view(torch.int64).sum().item())
        sums = [None] * self.world
        dist.all_gather_object(sums, my_sum, group=self.cpu_group)
        res = None
        if self.rank == 0:
            u0_h, p_h = inputs_np(cfg, 0, n_total, n_total)
            t_pin = time.perf_counter()
            u0_pin, p_pin = torch.from_numpy(u0_h).pin_memory(), torch.from_numpy(p_h).pin_memory()
            out_pin = torch.empty((3, n_total), dtype=torch.float64).pin_memory()
            t_pin = time.perf_counter() - t_pin
            keep = []
            alg, sysm = getattr(S, cfg["alg"])(), getattr(S.systems, cfg["system"])
            o = S.api.make_options(alg, np.dtype(np.float64), n_total, cfg["tspan"], cfg["dt"], 1e-6, 1e-3, None, _lib.SAVE_ENDPOINT,
                                   _lib.LAYOUT_TRAJ_MAJOR, 0, 0, keep)
            devs = (ctypes.c_int * self.world)(*range(self.world))

            def call():
                rc = _lib.lib().sde_solve(sysm._handle, ctypes.byref(o), u0_pin.data_ptr(), p_pin.data_ptr(), out_pin.data_ptr(), None,
                                          None, None, None, devs, self.world)
                _lib.check(rc)
            call()                       # warm-up: contexts, pools and streams on the other devices
            t0 = time.perf_counter()
            for _ in range(calls):
                call()
            secs = (time.perf_counter() - t0) / calls
            from simplediffeq_b200.sharding import shard_bounds
            ok = True
            for r in range(self.world):
                lo, hi = shard_bounds(n_total, self.world, r)
                ok = ok and int(out_pin[:, lo:hi].contiguous().view(torch.int64).sum().item()) == sums[r]
            res = {"api": "sde_solve(..., devices=[0..%d]) from ONE host process: one host thread + 2 streams per device inside the library" % (self.world - 1),
                   "workload": "%s, %d trajectories (%d per GPU) in pinned host memory" % (cfg["desc"], n_total, n_per_gpu),
                   "e2e_trajectory_steps_per_s": n_total * n_steps_of(cfg) / secs, "ms_per_call": secs * 1e3, "calls_timed": calls,
                   "h2d_bytes_per_call": int(n_total * 48), "d2h_bytes_per_call": int(n_total * 24), "pin_seconds": t_pin,
                   "bit_identical_to_per_rank_results": bool(ok)}
            del u0_pin, p_pin, out_pin
        dist.barrier(group=self.cpu_group)      # the other ranks wait on the CPU: their GPUs belong to rank 0 meanwhile
        return res


def cost_weighted(b, name, n_total):
    """Cost-weighted contiguous bounds for an adaptive sweep (simplediffeq_b200.sharding): every rank solves the same 4096
    pilot trajectories on its own device and inverts the cumulative attempt count -- no exchange."""
    from simplediffeq_b200.sharding import pilot_weighted_bounds
    torch, S = b.torch, b.S
    cfg = CONFIGS[name]
    sysm, alg = getattr(S.systems, cfg["system"]), getattr(S, cfg["alg"])()

    def pilot(idx):
        u0n, pn = inputs_np_idx(cfg, np.asarray(idx, dtype=np.int64), n_total)
        u0, p = torch.from_numpy(u0n).to(b.dev), torch.from_numpy(pn).to(b.dev)
        o = S.solve_device(sysm, alg, u0, p, cfg["tspan"], dt=DT0, abstol=cfg["tol"], reltol=cfg["tol"], compat=cfg.get("compat", 0))
        return (o["naccept"] + o["nreject"]).cpu().numpy()
    return pilot_weighted_bounds(n_total, b.world, pilot)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the config's ensemble per GPU, strong = the config's ensemble in total")
    ap.add_argument("--n-traj", type=int, default=None, help="override the config's ensemble size (per GPU if weak)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the selected config (no other configs / scaling studies)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    b = Bench(args)
    torch, S, _lib = b.torch, b.S, b._lib
    from simplediffeq_b200.sharding import shard_bounds, endpoint_stats, gather_endpoint_stats
    cfg = CONFIGS[args.config]
    n_cfg = args.n_traj or cfg["n"]
    if "saveat" in cfg and args.n_traj is None:      # 96 GB of output: only when the device has the room
        free, _ = torch.cuda.mem_get_info(b.dev)
        if args.scaling == "weak" and free < 110e9:
            n_cfg = 1_000_000
    n_total = n_cfg * b.world if args.scaling == "weak" else n_cfg
    lo, hi = shard_bounds(n_total, b.world, b.rank)
    warmup = max(args.warmup, 3)

    # roofline denominator measured on this device, now (dense DFMA / FFMA loop)
    peak_meas, _ = _lib.probe_fma_peak(_lib.SDE_F32 if cfg.get("dtype") == "f32" else _lib.SDE_F64)

    launch, st, keep = b.make_runner(cfg, lo, hi, n_total)
    launches0 = _lib.launch_count()
    for _ in range(warmup):
        launch()
    launches_warm = _lib.launch_count() - launches0
    ms_total, ms_kernel, clocks = b.time_launches(launch, 0, args.steps, sample_clocks=True)
    launches = _lib.launch_count() - launches0 - launches_warm
    acc, att = b.work_of(cfg, st, hi - lo)
    acc_all = b.reduce(acc, "sum")
    value = acc_all * args.steps / (ms_total * 1e-3)
    d_out = st["out"]["u"]

    e2e_steps = max(1, min(args.steps, 3))
    e2e = b.e2e_of(args.config, lo, hi, n_total, d_out, e2e_steps)
    stats = gather_endpoint_stats(endpoint_stats(d_out), b.dist) if d_out.dim() == 2 else None   # the only data collective, O(100 B)

    line = None
    if b.rank == 0:
        rl = b.roofline_of(cfg, acc / (ms_kernel * 1e-3), att / (ms_kernel * 1e-3), hi - lo, ms_kernel)
        if "flop" in cfg:       # the contract's roofline object for the compute-bound headline: flops against the FMA peak
            pipe = b.pipe_peak(cfg)
            achieved = acc / (ms_kernel * 1e-3) * cfg["flop"] / 1e12
            rl = {"bound": "fp64_fma" if cfg.get("dtype") != "f32" else "fp32_fma", "achieved": achieved, "peak": 2 * pipe / 1e12,
                  "unit": "TFLOP/s", "frac": achieved / (2 * pipe / 1e12),
                  "traffic": committed_traffic("r2_ncu_bench_traffic.json", hi - lo) if args.config == "2" else None,
                  "traffic_source": "committed ncu --set full capture of this launch (profiles/r2_ncu_bench_tsit5_10m.summary.txt), not a live measurement",
                  "traffic_unit": "bytes per launch (ncu dram read+write; algorithmic = %d)" % ((hi - lo) * bytes_per_traj(cfg)),
                  "peak_source": "derived 148 SM x %d FMA/clk x sm_max_mhz of MEASURED_PEAKS.json (no FP64 figure there); tensor cores n/a"
                                 % (FP32_LANES if cfg.get("dtype") == "f32" else FP64_LANES),
                  "peak_measured_fma_loop": peak_meas, "frac_of_measured_fma_loop": achieved / peak_meas if peak_meas else None,
                  "pipe_util": acc / (ms_kernel * 1e-3) * cfg["instr"] / pipe,
                  "kernel": "sde::fixed_kernel<Lorenz,%s,Tsit5Method,endpoint>" % ("float" if cfg.get("dtype") == "f32" else "double"),
                  "kernel_ms": ms_kernel, "flop_per_step": cfg["flop"], "pipe_instr_per_step": cfg["instr"],
                  "hbm_gbs": (hi - lo) * bytes_per_traj(cfg) / (ms_kernel * 1e-3) / 1e9, "hbm_peak_gbs": b.peaks.get("hbm_gbs")}
        else:
            rl["kernel_ms"] = ms_kernel
        line = {
            "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": b.world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": cfg.get("dtype", "f64"),
            "data": "synthetic",
            "config": {"workload": "BASELINE.json %s: %s%s" % (cfg["baseline"], cfg["desc"],
                                                               " -- %d trajectories per GPU" % n_cfg if args.scaling == "weak" and b.world > 1 else ""),
                       "name": args.config, "trajectories_total": n_total, "sharding": "contiguous index ranges, no data-path collective",
                       "l2": "inputs and outputs per pass (%d MB per GPU) exceed the 126 MB L2" % ((hi - lo) * bytes_per_traj(cfg) // 1000000)
                             if (hi - lo) * bytes_per_traj(cfg) > 126e6 else
                             "per-pass data (%d MB) fits L2; the kernel is FP64-issue bound (bytes per step negligible)" % ((hi - lo) * bytes_per_traj(cfg) // 1000000)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": rl,
        }
        if stats is not None:
            line["endpoint_stats"] = {"mean": [float(x) for x in stats["mean"]], "min": [float(x) for x in stats["min"]],
                                      "max": [float(x) for x in stats["max"]], "gathered_over_ranks": b.world}
        if is_adaptive(cfg):
            line["attempts_per_s"] = b.reduce(att, "sum") * args.steps / (ms_total * 1e-3)
    elif is_adaptive(cfg):
        b.reduce(att, "sum")

    # ---- CPU baseline, parity evidence (rank 0, N = 1 only)
    if b.rank == 0 and b.world == 1 and not args.no_cpu_baseline:
        import oracle_lib
        oracle_lib.build()
        cores = oracle_lib.hardware_threads()
        n_s = min(cfg["n"], cpu_sample_size(cfg, cores))
        acc_c, secs, _ = cpu_oracle_run(cfg, n_s, cores)
        line["cpu_baseline"] = {"value": acc_c / secs, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                                "sample": "%d of %d trajectories in %.1f s, %d std::threads (C++ restatement of %s; Julia unavailable)"
                                          % (n_s, cfg["n"], secs, cores, cfg["alg"])}
        # parity of what was just computed: a slice of THIS sweep through the oracle
        m = 512
        u0c, pc = inputs_np(cfg, 0, m, n_total)
        from simplediffeq_b200 import jl_range
        if is_adaptive(cfg):
            r = oracle_lib.solve(cfg["system"], cfg["oalg"], u0c.T, pc.T, cfg["tspan"][0], cfg["tspan"][1], DT0, abstol=cfg["tol"],
                                 reltol=cfg["tol"], dtype=np_dtype(cfg), n_threads=cores)
            g = st["out"]
            line["parity_spot_check"] = {"trajectories": m,
                                         "identical_step_counts_frac": float(np.mean(g["naccept"][:m].cpu().numpy() == r.naccept)),
                                         "bit_identical_final_state_frac": float(np.mean(np.all(
                                             g["u"][:, :m].cpu().numpy().T.view(np.uint64) == np.ascontiguousarray(r.u[:, 0, :]).view(np.uint64), axis=1)))}
        elif d_out.dim() == 2:
            dt_ = np_dtype(cfg)
            r = oracle_lib.solve(cfg["system"], cfg["oalg"], u0c.T, pc.T, cfg["tspan"][0], cfg["tspan"][1], cfg["dt"], dtype=dt_,
                                 tgrid=jl_range(dt_(cfg["tspan"][0]), dt_(cfg["dt"]), dt_(cfg["tspan"][1]), dt_), n_threads=cores)
            line["parity_spot_check"] = {"trajectories": m,
                                         "bit_identical_to_oracle": bool(np.array_equal(r.u[:, 0, :].T, d_out[:, :m].cpu().numpy()))}

    # ---- everything else the metric names, on the same device(s): other configs, scaling studies, in-library sharder
    if not args.no_extras and args.config == "2":
        d_keep = d_out          # per-rank result of the headline workload (for the in-library comparison)
        extras = {}
        study = {}
        try:
            del launch, st, keep
            for name in EXTRA_ORDER:
                c = CONFIGS[name]
                n_c = c["n"]
                if "saveat" in c:
                    free, _ = torch.cuda.mem_get_info(b.dev)
                    per_gpu = n_c // b.world
                    if per_gpu * bytes_per_traj(c) > 0.85 * free:
                        n_c = int(0.8 * free / bytes_per_traj(c)) // 1024 * 1024 * b.world
                try:
                    r = b.run_config(name, n_total=n_c, steps=5 if "saveat" not in c else 3)
                    if b.world > 1 and is_adaptive(c) and n_c >= 4096 * b.world:
                        # adaptive sweeps: step counts vary along the index, so equal index ranges are not equal work.  The
                        # entry reports the cost-weighted contiguous split; the equal split stays beside it.
                        b.free()
                        eq = {k: r[k] for k in ("ms", "trajectory_steps_per_s", "attempts_per_s", "shard_sizes")}
                        r = b.run_config(name, n_total=n_c, bounds=cost_weighted(b, name, n_c), steps=5,
                                         split="cost-weighted contiguous ranges (sharding.pilot_weighted_bounds)")
                        r["equal_index_ranges"] = eq
                    if b.world == 1 and not args.no_cpu_baseline and b.rank == 0 and is_adaptive(c) and name in ("1", "4"):
                        r.update(adaptive_parity(b, name))
                    extras[name] = r
                except Exception as e:          # e.g. not enough free memory next to another tenant: reported, never fatal
                    extras[name] = {"error": str(e)[:300]}
                b.free()
            if b.world > 1:
                # strong scaling of the headline config: the config's own 10 M trajectories in total
                study["config2_strong"] = b.run_config("2", n_total=CONFIGS["2"]["n"], steps=max(5, min(args.steps, 20)))
                b.free()
                # weak scaling of the HBM-bound config: 4 M trajectories (96 GB of output) per GPU
                free, _ = torch.cuda.mem_get_info(b.dev)
                if free > 110e9:
                    study["config5_weak"] = b.run_config("5", n_total=CONFIGS["5"]["n"] * b.world, steps=5)
                    b.free()
                # adaptive sweeps at a GPU-filling size per device (2^20 per GPU), equal index ranges vs the cost-weighted split
                for name in ("3", "3s", "1b"):
                    n_w = (1 << 20) * b.world
                    study["config%s_weak_equal" % name] = b.run_config(name, n_total=n_w, steps=5)
                    b.free()
                    bounds = cost_weighted(b, name, n_w)
                    study["config%s_weak_cost_weighted" % name] = b.run_config(name, n_total=n_w, bounds=bounds, steps=5,
                                                                               split="cost-weighted contiguous ranges (sharding.pilot_weighted_bounds)")
                    b.free()
                n_w = CONFIGS["4"]["n"] * b.world
                study["config4_weak_equal"] = b.run_config("4", n_total=n_w, steps=5)
                b.free()
                study["config4_weak_cost_weighted"] = b.run_config("4", n_total=n_w, bounds=cost_weighted(b, "4", n_w), steps=5,
                                                                   split="cost-weighted contiguous ranges (sharding.pilot_weighted_bounds)")
                b.free()
                study["inlib_sharder"] = b.inlib_sharder("2", n_cfg, d_keep)
        except Exception as e:
            study["error"] = str(e)[:300]
        if b.rank == 0:
            line["configs"] = extras
            if study:
                line["multi_gpu"] = study
    if b.rank == 0:
        print(json.dumps(line), flush=True)
    if b.dist is not None:
        b.dist.destroy_process_group()


def adaptive_parity(b, name, m=4096):
    """Outside every timed region: the first m trajectories of an adaptive config through the public host API with DEFAULT
    options against the CPU oracle -- share of identical accepted AND rejected counts, share of bit-identical final states."""
    import oracle_lib
    S = b.S
    cfg = CONFIGS[name]
    m = cfg["n"] if cfg["n"] <= 10_000 else min(m, cfg["n"])
    u0, p = inputs_np(cfg, 0, m, cfg["n"])
    cores = oracle_lib.hardware_threads()
    t0 = time.perf_counter()
    o = oracle_lib.solve(cfg["system"], cfg["oalg"], u0.T, p.T, cfg["tspan"][0], cfg["tspan"][1], DT0, abstol=cfg["tol"], reltol=cfg["tol"],
                         n_threads=cores)
    cpu_s = time.perf_counter() - t0
    e2e_ms = None
    for rep in range(4):        # first call untimed (module load, pool growth), then the best of three
        t0 = time.perf_counter()
        g = S.solve_arrays(getattr(S.systems, cfg["system"]), getattr(S, cfg["alg"])(), u0, p, cfg["tspan"], dt=DT0, abstol=cfg["tol"],
                           reltol=cfg["tol"], compat=cfg.get("compat", 0), devices=[b.local])
        ms = (time.perf_counter() - t0) * 1e3
        if rep > 0:
            e2e_ms = ms if e2e_ms is None else min(e2e_ms, ms)
    gu, ou = np.ascontiguousarray(g["u"].T), np.ascontiguousarray(o.u[:, 0, :])
    return {"parity_vs_oracle": {"trajectories": m, "options": "default",
                                 "identical_step_counts_frac": float(np.mean((g["naccept"] == o.naccept) & (g["nreject"] == o.nreject))),
                                 "bit_identical_final_state_frac": float(np.mean(np.all(gu.view(np.uint64) == ou.view(np.uint64), axis=1))),
                                 "host_api_ms": e2e_ms, "host_api_timing": "solve_arrays (host numpy buffers, pageable) wall clock, best of 3 after one untimed call", "host_api_steps_per_s": int(g["naccept"].sum()) / (e2e_ms * 1e-3),
                                 "cpu_oracle_steps_per_s": int(o.naccept.sum()) / cpu_s, "cpu_cores": cores}}


if __name__ == "__main__":
    main()
