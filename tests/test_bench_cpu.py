"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the required keys, and the
roofline `traffic` figure is read from the committed ncu capture."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("force_port", [False, True])
def test_reference_arm_prints_one_json_line(force_port):
    """Unmodified reference modules where /root/reference (or baseline/_ref) exists, the oracle port otherwise / when forced."""
    env = dict(os.environ)
    if force_port:
        env["KG_REFERENCE_PORT"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    have_ref = os.path.exists("/root/reference/KGnet.py") or os.path.exists(os.path.join(ROOT, "baseline", "_ref", "KGnet.py"))
    assert cb["kind"] == ("reference" if have_ref and not force_port else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"]


def test_roofline_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    per_launch, src = bench.profiled_traffic()
    assert src is not None and "_traffic.json" in src and per_launch > 1e8
    blur, _ = bench.profiled_traffic("decode")
    assert 5e7 < blur < 5e8        # ~446 MB of accumulators per step over four launches


def test_bench_has_no_undefined_names():
    """The GPU legs of bench.py cannot run here: at least every name their (nested) functions read must resolve to a module-level
    definition, an import or a builtin (a stray line once left `drain()` in the decode leg)."""
    import builtins
    import symtable
    path = os.path.join(ROOT, "bench.py")
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    missing = []

    def walk(tab):
        for s in tab.get_symbols():
            if s.is_global() and s.is_referenced() and not s.is_assigned():
                n = s.get_name()
                if n not in module_names and not hasattr(builtins, n):
                    missing.append((tab.get_name(), n))
        for ch in tab.get_children():
            walk(ch)

    for ch in top.get_children():
        walk(ch)
    assert not missing, missing
