"""bench.py output contract (what the round driver parses): the reference arm is executed here on the CPU for one
bounded step; the committed GPU line under profiles/ is checked for the same keys plus the roofline / cpu_baseline /
e2e / clocks objects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _baseline():
    return json.load(open(os.path.join(ROOT, "BASELINE.json")))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["unit"] == "patches/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["config"]["batch_per_gpu"] == 2 and line["cpu_processes"] == 1      # same per-step batch as the B200 arm
    assert abs(line["value"] - 2 / line["ms_per_step"] * 1e3) < 1e-6 * line["value"]
    assert line["metric"].split(" at ")[0] in _baseline()["metric"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                          "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r01_bench_n1.json")
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(line)
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["dtype"] == "f32"
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    c = line["clocks"]
    assert c["sm_max_mhz"] >= c["sm_mhz"] > 0
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0
    assert abs(line["value"] - line["config"]["global_batch"] / line["ms_per_step"] * 1e3) < 1e-6 * line["value"]
