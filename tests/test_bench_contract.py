"""bench.py's output contract, checked on the committed bench lines (profiles/) and on the argument surface -- no GPU."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}


def _lines():
    return sorted(glob.glob(os.path.join(ROOT, "profiles", "r01b_bench_c*_n*.json")))


def test_committed_bench_lines_follow_the_contract():
    files = _lines()
    assert files, "no committed bench lines"
    for f in files:
        j = json.load(open(f))
        assert BASE_KEYS <= set(j), (f, BASE_KEYS - set(j))
        assert j["metric"] == "velocity_fields_per_sec_fwd_bwd" and j["unit"] == "fields/s" and j["higher_is_better"] is True
        assert j["scaling"] == "weak" and j["vs_baseline"] is None and j["data"] == "synthetic"
        assert j["warmup"] >= 3 and j["gpu_launches"] > 0
        assert abs(j["value"] - j["config"]["global_batch"] / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
        assert "workload" in j["config"] and "model" not in j["config"]
        e = j["e2e"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        r = j["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        c = j["clocks"]
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])), (f, c)
        if j["n_gpus"] == 1 and j["cpu_baseline"] is not None:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"]) and j["cpu_baseline"]["kind"] in ("port", "reference")


def test_reference_arm_line():
    j = json.load(open(os.path.join(ROOT, "profiles", "r01b_bench_reference_arm_c4.json")))
    assert j["impl"] == "reference" and j["metric"] == "velocity_fields_per_sec_fwd_bwd" and j["unit"] == "fields/s"
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["cores"] >= 1


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--precision"):
        assert flag in out.stdout


def test_vendor_bar_tool_runs_tiny_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cudnn_bar.py"), "--device", "cpu", "--tiny", "--workload", "c2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j["impl"] == "torch-cpu" and j["value"] > 0 and j["params"] > 0
