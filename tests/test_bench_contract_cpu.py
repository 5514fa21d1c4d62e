"""The committed bench lines (profiles/r1_bench_*.json, written by bench.py on the B200 boxes) carry every key of the
driver's contract; bench.py's own argument defaults match it.  No GPU, no reference, no oracle needed."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r1_bench_n1.json", "r1_bench_n2.json", "r1_bench_n4.json", "r1_bench_n8.json"])
def test_bench_line_has_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["metric"] == "FF+Sinkhorn clips/s" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # whole-job throughput = clips of all ranks / step time
    assert abs(d["value"] - d["config"]["global_clips"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]


def test_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "FF+Sinkhorn clips/s" and d["unit"] == "clips/s"
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_bench_defaults_finish_in_minutes():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert re.search(r'"--gpus", type=int, default=1\b', src)
    steps = int(re.search(r'"--steps", type=int, default=(\d+)', src).group(1))
    warm = int(re.search(r'"--warmup", type=int, default=(\d+)', src).group(1))
    assert warm >= 3 and steps <= 100
