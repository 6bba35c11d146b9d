"""Host-side logic of bench.py that needs no GPU: the synthetic inputs of the BASELINE configs, the
algorithmic byte counts behind the HBM roofline, the config table against BASELINE.json, and the CPU reference arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_inputs_follow_the_survey_formulas():
    n = 10_000
    u0, p = bench.inputs_np(bench.CONFIGS["1"], 0, n, n)
    assert u0.shape == (3, n) and p.shape == (3, n) and u0.dtype == np.float64
    assert np.all(u0[0] == 1) and np.all(u0[1:] == 0) and np.all(p[0] == 10) and np.all(p[2] == 8.0 / 3.0)
    i = np.arange(n)
    assert np.array_equal(p[1], (21.0 * i) / float(n - 1))          # rho_i = double(21 i) / double(N - 1)
    assert p[1][0] == 0.0 and p[1][-1] == 21.0
    # a shard is a slice of the whole sweep
    a, b = bench.inputs_np(bench.CONFIGS["2"], 1234, 2345, 10_000_000)
    assert np.array_equal(b[1], (21.0 * np.arange(1234, 2345)) / 9_999_999.0)
    # Float32: every literal rounded once from double; 8/3f0
    u32, p32 = bench.inputs_np(bench.CONFIGS["2f32"], 0, 100, 100)
    assert p32.dtype == np.float32 and p32[2][0] == np.float32(8.0) / np.float32(3.0)
    assert np.array_equal(p32[1], ((21.0 * np.arange(100)) / 99.0).astype(np.float32))
    # Van der Pol: mu_i = 0.1 + 49.9 i / (N - 1); shuffled = the same multiset, i -> i * 2654435761 mod N (a bijection)
    m = 1 << 20
    v0, mu = bench.inputs_np(bench.CONFIGS["3"], 0, m, m)
    assert v0.shape == (2, m) and np.all(v0[0] == 2) and np.all(v0[1] == 0)
    assert mu[0][0] == 0.1 and abs(mu[0][-1] - 50.0) < 1e-12 and np.all(np.diff(mu[0]) > 0)
    _, mus = bench.inputs_np(bench.CONFIGS["3s"], 0, m, m)
    assert not np.array_equal(mus, mu) and np.array_equal(np.sort(mus[0]), mu[0])
    k = np.arange(5, dtype=np.int64)
    assert np.array_equal(mus[0][:5], mu[0][(k * 2654435761) % m])


def test_config_table_matches_baseline_json():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    c = bench.CONFIGS
    assert c["1"]["n"] == 10_000 and c["1"]["tol"] == 1e-8 and c["1"]["alg"] == "GPUSimpleATsit5"
    assert c["2"]["n"] == 10_000_000 and c["2"]["dt"] == 1e-3 and bench.n_steps_of(c["2"]) == 10_000
    assert c["3"]["n"] == 1 << 20 and c["3"]["tol"] == 1e-6 and c["3"]["system"] == "vanderpol" and c["3"]["tspan"] == (0.0, 20.0)
    assert c["4"]["n"] == 1_000_000 and c["4"]["tol"] == 1e-12 and c["4"]["alg"] == "GPUSimpleAVern9" and "compat" not in c["4"]
    assert c["5"]["n"] == 4_000_000 and c["5"]["saveat"] == (0.0, 0.01, 10.0)
    # algorithmic bytes per trajectory: 72 B endpoint FP64 (36 FP32), 24 072 B with 1001 save points, 40 B Van der Pol
    assert bench.bytes_per_traj(c["2"]) == 72 and bench.bytes_per_traj(c["2f32"]) == 36
    assert bench.bytes_per_traj(c["5"]) == 24_072 == bench.bytes_per_traj(c["5tm"])
    assert bench.bytes_per_traj(c["3"]) == 40
    assert bench.n_steps_of(c["5"]) == 100 and bench.n_steps_of(c["5f"]) == 1000
    for name in bench.EXTRA_ORDER:
        assert name in c and name != "2"
    for name, cfg in c.items():
        assert bench.is_adaptive(cfg) == (cfg["alg"] in ("GPUSimpleATsit5", "GPUSimpleAVern7", "GPUSimpleAVern9")), name
        assert 0 < bench.cpu_sample_size(cfg, 8) <= cfg["n"]


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` on the CPU (the oracle port), default config and an adaptive one: one JSON line with the
    contract's keys; ranks other than 0 print nothing."""
    for cfg in ("2", "1"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                              "--config", cfg], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-500:]
        lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1
        line = json.loads(lines[0])
        assert line["impl"] == "reference" and line["unit"] == "trajectory-steps/s" and line["value"] > 1e6
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
        assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0 and line["higher_is_better"] is True
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
