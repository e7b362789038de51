"""The JSON line bench.py prints is a contract with the driver: check its keys on both arms.
The reference arm runs on the CPU (oracle port, bounded sample); the B200 arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--workload", "hgp_34_n225", "--steps", "1", "--warmup", "1", "--cpu-seconds", "0.3")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "shots/s" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert isinstance(d["ms_per_step"], float) and d["ms_per_step"] > 0      # wall time of one step's bounded sample
    assert d["e2e"] == {"value": d["value"], "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "shots" in cb["sample"]
    assert d["config"]["workload"].startswith("hgp_34_n225") and "model" not in d["config"]
    one = cb["single_core"]                                                  # all cores and one core
    assert one["unit"] == "shots/s" and one["value"] > 0 and "1 threads" in one["sample"] and "upper bound" in cb["note"]


def test_reference_arm_ssf_runs_without_the_bp_round_cap():
    """The SSF baseline must do the GPU decoder's work: no round cap (it once inherited BP's max_iter = 32)."""
    d = _run("--impl", "reference", "--workload", "hgp_34_n225", "--decoder", "ssf", "--steps", "1", "--warmup", "1",
             "--cpu-seconds", "0.3")
    assert d["impl"] == "reference" and "SSF" in d["metric"] and d["dtype"] == "u32"
    assert d["config"]["decoder"] == "ssf" and d["config"]["max_iter"] == 0 and d["value"] > 0


@pytest.mark.gpu
def test_b200_arm_line(built):
    d = _run("--workload", "hgp_34_n225", "--shots", "65536", "--steps", "2", "--warmup", "3", "--cpu-seconds", "0.5")
    assert BASE | {"gpu_launches", "clocks", "roofline", "roofline_smem"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["value"] > 0 and d["gpu_launches"] == 2
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] == "bp7_kernel" and r["kernel_ms"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert d["counters"]["shots"] == 2 * 65536
