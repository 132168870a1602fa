"""bench.py's GPU legs end to end on the smallest configuration (BASELINE config 1: bs1, 512 x 512): every mode the driver or
tools/gpu_profile.sh runs must print exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                        "--no-gpu-baseline", *flags], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("flags", [("--config", "cfg1"), ("--config", "cfg1", "--serial"), ("--config", "cfg1", "--free-running")])
def test_pipeline_line(flags):
    d = _run(*flags)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 512 * 512 * 3
    assert d["gpu_launches"] > 0 and d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1.5
    assert d["config"]["batches_in_flight"] == (1 if "--serial" in flags else 2)
    if "--free-running" not in flags:
        assert d["config"]["detections_per_step"] > 0


def test_decode_line():
    d = _run("--workload", "decode")
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
