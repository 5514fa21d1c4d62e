"""bench.py on the CPU: the reference arm is really run (small config) and its JSON line parsed; the workload resolution
(weak / strong scaling) and the rank gating of the reference arm are checked.  The GPU arm needs a
B200 and is exercised by the driver."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("impl", ["port", "auto"])
def test_reference_arm_runs_and_prints_the_contract_line(impl):
    r = _run(["--impl", "reference", "--config", "1", "--steps", "2", "--warmup", "1", "--cpu-impl", impl])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]
    if impl == "port":
        assert cb["kind"] == "port"
    assert "configs[0]" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--config", "1", "--steps", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_resolution():
    ns = lambda **kw: types.SimpleNamespace(**{"config": 2, "scaling": None, "clips": None, **kw})
    c = bench.resolve(ns(), 8)
    assert (c["clips_per_gpu"], c["global_clips"], c["scaling"]) == (32, 256, "weak")
    c = bench.resolve(ns(config=3), 4)                       # BASELINE configs[2]: 256 clips GLOBAL
    assert (c["clips_per_gpu"], c["global_clips"], c["scaling"]) == (64, 256, "strong")
    c = bench.resolve(ns(config=3), 1)
    assert c["clips_per_gpu"] == 256
    c = bench.resolve(ns(config=5, clips=4), 8)
    assert (c["clips_per_gpu"], c["global_clips"], c["sr"], c["dim"], c["K"], c["fs"]) == (4, 32, 56, 768, 300, 16)
    with pytest.raises(SystemExit):
        bench.resolve(ns(config=3, clips=10), 4)
    assert bench.CONFIGS[4]["radius"] == 12 and bench.CONFIGS[4]["topk"] == 7 and bench.CONFIGS[4]["fs"] == 80
    assert bench.parse.__defaults__ is None                   # defaults: N = 1, configs[1], finishes within minutes
    a = bench.parse.__globals__["argparse"].ArgumentParser
    assert a is not None


def test_defaults_are_the_headline_config(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.config, a.impl, a.steps >= 10, a.warmup >= 3) == (1, 2, "ours", True, True)
