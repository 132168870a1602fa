"""Drop-in for the reference's nms.py.  The arithmetic runs in the CUDA `nms_kernel`
(csrc/decode.cu) through the C-ABI entry point kg_nms_host."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


def non_maximum_suppression_numpy(bboxes, nms_thresh=0.5):
    """nms.py:4-53: bboxes (n,5) rows [y1,x1,y2,x2,conf] -> kept rows in descending-conf keep order,
    or None when the input is empty (the reference's convention, test.py:119)."""
    if len(bboxes) == 0:
        return None
    b = np.ascontiguousarray(np.asarray(bboxes, np.float64).reshape(-1, 5))
    out = np.empty_like(b)
    n_out = C.c_int(0)
    _cabi.check(_cabi.lib().kg_nms_host(b.ctypes.data, len(b), float(nms_thresh), out.ctypes.data, C.byref(n_out)))
    return out[:n_out.value].copy()
