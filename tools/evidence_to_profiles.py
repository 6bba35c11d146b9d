"""Copy the outputs of tools/r2_evidence.sh from gpurun_out/ into profiles/ and derive the small files other things read:
  r2_launches_bench_summary.txt   launch list aggregated by kernel (shares of the bench run)
  r2_ncu_*_traffic.json           DRAM bytes per launch from the ncu summaries (bench.py's roofline.traffic)
CPU only; run after the evidence call has been merged back."""
import csv, json, os, re, shutil, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

COPY = ["r2_gpu_tests_final.txt", "r2_bench_n1_final.json", "r2_bench_reference_arm.json", "r2_launches_bench.csv"]
for f in sorted(os.listdir(SRC)):
    if re.match(r"r2_ncu_.*\.(summary|regions|opcodes)\.txt$", f):
        COPY.append(f)
for f in COPY:
    if os.path.exists(os.path.join(SRC, f)):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))

# ---- launch list -> per-kernel shares ------------------------------------------------------------------------------
rows = []
with open(os.path.join(SRC, "r2_launches_bench.csv")) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        rows.append((r["Kernel Name"], ms))
agg = defaultdict(lambda: [0, 0.0])
for k, ms in rows:
    agg[k][0] += 1
    agg[k][1] += ms
total = sum(v[1] for v in agg.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` on one B200 (round 2, final tree).",
       "Every launch of the run (headline config 2 + e2e + the other configs of the default line); per-launch times are cold-cache and serialised: SHARES, not absolutes.",
       "%-112s %6s %12s %7s" % ("kernel", "count", "total ms", "share")]
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-112s %6d %12.3f %6.2f%%" % (k[:110], c, ms, 100 * ms / total))
out.append("total %.1f ms in %d launches; kernels of this repository only (libsimplediffeq_cuda.so) plus torch's fill/copy helpers for the inputs" % (total, len(rows)))
open(os.path.join(DST, "r2_launches_bench_summary.txt"), "w").write("\n".join(out) + "\n")


# ---- ncu summaries -> traffic JSONs --------------------------------------------------------------------------------
def metric(text, name):
    m = re.search(r"^\s*%s\s+([0-9.]+)\s*(\S*)" % re.escape(name), text, re.M)
    if not m:
        return None
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
    return float(m.group(1)) * scale.get(m.group(2), 1)


def traffic(summary, dst, **meta):
    path = os.path.join(SRC, summary)
    if not os.path.exists(path):
        return
    t = open(path).read()
    d = dict(meta)
    d["dram_bytes_read"] = int(metric(t, "dram__bytes_read.sum"))
    d["dram_bytes_write"] = int(metric(t, "dram__bytes_write.sum"))
    d["gpu_time_duration_ns"] = int(metric(t, "gpu__time_duration.sum"))
    json.dump(d, open(os.path.join(DST, dst), "w"), indent=1)
    print(dst, d["dram_bytes_read"], d["dram_bytes_write"], d["gpu_time_duration_ns"])


traffic("r2_ncu_bench_tsit5_10m.summary.txt", "r2_ncu_bench_traffic.json",
        kernel="sde::fixed_kernel<Lorenz,double,Tsit5Method,endpoint>", n_traj=10000000, n_steps=10000, algorithmic_bytes=720000000,
        source="ncu --set full --clock-control none, one launch of `python bench.py --steps 1 --warmup 3 --no-extras` (profiles/r2_ncu_bench_tsit5_10m.summary.txt)")
traffic("r2_ncu_config5_soa.summary.txt", "r2_ncu_config5_traffic.json",
        kernel="sde::fixed_kernel<Lorenz,double,Tsit5Method,saveat,SoA>", n_traj=4000000, n_save=1001, dt=0.1, algorithmic_bytes=96288000000,
        source="ncu --set full --clock-control none, one launch of `python tools/prof_saveat.py 1 4000000 0.1` (profiles/r2_ncu_config5_soa.summary.txt)")
traffic("r2_ncu_config5_tm.summary.txt", "r2_ncu_config5tm_traffic.json",
        kernel="sde::fixed_kernel<Lorenz,double,Tsit5Method,saveat,traj-major staged (whole-line flushes)>", n_traj=4000000, n_save=1001, dt=0.1,
        algorithmic_bytes=96288000000,
        source="ncu --set full --clock-control none, one launch of `python tools/prof_saveat.py 0 4000000 0.1` (profiles/r2_ncu_config5_tm.summary.txt)")
