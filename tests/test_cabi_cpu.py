"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/*.h declares;
host constants baked into the kernels equal what SciPy/NumPy compute."""
import ctypes
import glob
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(kg_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from kg_instance_segmentation_b200 import build, _cabi
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert _cabi.lib().kg_abi_version() == 2
    for s in _cabi.EXPORTS:
        assert s in syms


def test_gaussian_taps_and_constants_match_scipy():
    from scipy.ndimage import _filters
    w = _filters._gaussian_kernel1d(2.0, 0, 8)
    src = open(os.path.join(ROOT, "kg_instance_segmentation_b200", "csrc", "decode.cu")).read()
    m = re.search(r"c_gauss\[9\]\s*=\s*\{([^}]*)\}", src)
    taps = [float.fromhex(t.strip()) for t in m.group(1).split(",")]
    assert taps == list(w[:9]) and list(w[9:]) == list(w[:8][::-1])
    m = re.search(r"KG_PI_R2\s*=\s*([0-9a-fx.p+-]+);", src)
    assert float.fromhex(m.group(1)) == np.pi * 5 ** 2


def test_mid_index_table_matches_config():
    from kg_instance_segmentation_b200 import config as cfg
    src = open(os.path.join(ROOT, "kg_instance_segmentation_b200", "csrc", "decode.cu")).read()
    m = re.search(r"c_mid_index\[5\]\[5\]\s*=\s*\{(.*?)\};", src, flags=re.S)
    vals = [int(v) for v in re.findall(r"-?\d+", m.group(1))]
    dir_edges = cfg.EDGES + [e[::-1] for e in cfg.EDGES]
    for s in range(5):
        for t in range(5):
            assert vals[s * 5 + t] == (-1 if s == t else dir_edges.index((s, t)))


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle or call torch conv ops."""
    pkg = os.path.join(ROOT, "kg_instance_segmentation_b200")
    for f in glob.glob(os.path.join(pkg, "*.py")):
        src = open(f).read()
        assert "oracle" not in src.replace("no CPU or PyTorch-op fallback", ""), f
        assert "F.conv2d" not in src and "scipy" not in src, f


def test_workspace_placement_never_overlaps_live_buffers():
    """kg_debug_place_by_liveness = the first-fit placement that lays out forward_dec's activation workspace.  Property: two buffers
    whose lifetimes [def, last] intersect -- or touch: an op must not write into the memory of its own inputs -- never share a byte;
    persistent buffers (last >= n_ops) are never reused; the arena is no larger than the consecutive layout."""
    import ctypes as C
    import numpy as np
    from kg_instance_segmentation_b200 import _cabi
    L = _cabi.lib()
    rs = np.random.RandomState(0)
    for trial in range(200):
        n_ops = int(rs.randint(1, 60))
        nb = int(rs.randint(0, 80))
        d = rs.randint(0, n_ops, nb).astype(np.int32)
        span = rs.randint(0, 12, nb)
        last = np.minimum(d + span, n_ops - 1).astype(np.int32)
        keep = rs.rand(nb) < 0.1
        last[keep] = n_ops + 1
        size = (rs.randint(0, 50, nb) * 1024).astype(np.uint64)        # zero-sized buffers occur (tensors without a lo plane)
        off = np.zeros(nb, np.uint64)
        total = C.c_ulonglong(0)
        _cabi.check(L.kg_debug_place_by_liveness(n_ops, nb, d.ctypes.data, last.ctypes.data, size.ctypes.data, off.ctypes.data,
                                                 C.byref(total)))
        assert total.value <= int(size.sum())
        for a in range(nb):
            assert int(off[a]) + int(size[a]) <= total.value
            for b in range(a + 1, nb):
                if size[a] == 0 or size[b] == 0:
                    continue
                live_together = not (last[a] < d[b] or last[b] < d[a])
                if live_together:
                    assert int(off[a]) + int(size[a]) <= int(off[b]) or int(off[b]) + int(size[b]) <= int(off[a]), (trial, a, b)
    # a chain a -> b -> c of equal sizes needs two slots, not three
    d = np.array([0, 1, 2], np.int32); last = np.array([1, 2, 2], np.int32); size = np.array([4096] * 3, np.uint64)
    off = np.zeros(3, np.uint64); total = C.c_ulonglong(0)
    _cabi.check(L.kg_debug_place_by_liveness(3, 3, d.ctypes.data, last.ctypes.data, size.ctypes.data, off.ctypes.data, C.byref(total)))
    assert total.value == 8192 and off[2] == off[0]
    assert L.kg_debug_place_by_liveness(3, 1, np.array([5], np.int32).ctypes.data, last.ctypes.data, size.ctypes.data, off.ctypes.data,
                                        C.byref(total)) != 0
