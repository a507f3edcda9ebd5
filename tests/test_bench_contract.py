"""bench.py prints ONE JSON line with the keys the measurement contract names (small sizes; the numbers themselves are
not checked here)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-rows", "1000000", "--uniques", "1000"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "rows/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the CPU sample is fixed-size: asking for more steps must not shrink it
    d5 = _run(["--impl", "reference", "--steps", "5", "--warmup", "1", "--cpu-rows", "1000000", "--uniques", "1000"], 300)
    assert d["config"]["rows_per_step"] == d5["config"]["rows_per_step"] == 1_000_000 == d5["cpu_baseline"]["rows"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--rows", "8000000", "--uniques", "10000", "--steps", "2", "--warmup", "3", "--cpu-rows", "1000000"], 600)
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["cpu_baseline"]["parity"].startswith("ok"), d["cpu_baseline"]["parity"]
    assert d["two_call"]["value"] > 0 and d["config"]["step"].startswith("one rc.MultiKeyGroupBySum32")
