"""The measurement contract of bench.py: the reference arm runs on the host cores (it times the CPU oracle port, the
one place besides tests/ and smoke() that may execute oracle/), and the committed B200 line carries every key the
contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "env-steps/s" and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_committed_b200_line_has_every_contract_key():
    with open(os.path.join(ROOT, "profiles", "r01_bench_default_1gpu.json")) as f:
        d = json.loads(f.read())
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["unit"] == base.get("unit", d["unit"]) and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["bytes_per_env_step"] * d["config"]["global_batch"] / (r["kernel_ms"] / 1e3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["value"] != d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs the CPU port and prints the line; the other ranks exit 0 without work."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29583", os.path.join(ROOT, "bench.py"),
                          "--gpus", "2", "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["value"] == d["value"]
