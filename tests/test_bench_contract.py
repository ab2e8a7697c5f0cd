"""bench.py's output contract (the driver parses it): checked on the committed round-1 lines under profiles/ and on the
reference arm, which runs on the host alone."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"]


def _line(path):
    return json.loads([ln for ln in open(path).read().splitlines() if ln.strip().startswith("{")][-1])


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r01_bench_n2.json", "r01_bench_n4.json", "r01_bench_n8.json"])
def test_committed_bench_lines_follow_the_contract(name):
    b = _line(os.path.join(ROOT, "profiles", name))
    for k in REQUIRED:
        assert k in b, k
    assert b["unit"] == "Gcell/s" and b["higher_is_better"] is True and b["scaling"] == "weak" and b["dtype"] == "f64"
    assert b["vs_baseline"] is None and "workload" in b["config"] and "model" not in b["config"]
    r = b["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    e = b["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < b["value"]
    assert b["gpu_launches"] > 0 and b["warmup"] >= 3
    c = b["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if b["n_gpus"] == 1:
        cb = b["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0
        kernels = {row["kernel"] for row in b["suite"] if "ms" in row}
        assert {"jacobi_2d", "heat_3d", "fdtd_2d", "hdiff", "vadv"} <= kernels


def test_reference_arm_runs_on_the_host(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    b = json.loads([ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")][-1])
    assert b["impl"] == "reference" and b["value"] > 0 and b["unit"] == "Gcell/s"
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["cpu_baseline"]["kind"] in ("port", "reference") and b["cpu_baseline"]["value"] == b["value"]
