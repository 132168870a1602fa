"""ctypes binding of libkgnet_b200.so (include/kgnet_b200.h).  There is NO fallback: if the library is
missing or a call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libkgnet_b200.so")

KG_MAX_SCALES = 4


class KgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kgnet_b200 error {code}: {msg}")
        self.code = code


class DecodeScale(C.Structure):
    _fields_ = [("d_kp", C.c_void_p), ("d_short", C.c_void_p), ("d_mid", C.c_void_p),
                ("H", C.c_int), ("W", C.c_int), ("box_scale", C.c_int)]


class DecodeConfig(C.Structure):
    _fields_ = [("N", C.c_int), ("n_scales", C.c_int), ("max_peaks", C.c_int), ("max_boxes", C.c_int),
                ("nms_thresh", C.c_double), ("peak_thresh", C.c_double)]


class DecodeOutputs(C.Structure):
    _fields_ = [("d_dets", C.c_void_p), ("d_det_count", C.c_void_p), ("d_boxes", C.c_void_p), ("d_box_count", C.c_void_p),
                ("d_skeletons", C.c_void_p), ("d_skel_count", C.c_void_p), ("d_skel_keep", C.c_void_p),
                ("d_peak_conf", C.c_void_p), ("d_peak_key", C.c_void_p), ("d_peak_count", C.c_void_p),
                ("d_heat", C.c_void_p * KG_MAX_SCALES), ("d_vote", C.c_void_p * KG_MAX_SCALES),
                ("d_status", C.c_void_p), ("d_det_packed", C.c_void_p), ("det_packed_k", C.c_int)]


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m kg_instance_segmentation_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.kg_last_error.restype = C.c_char_p
    L.kg_abi_version.restype = C.c_int
    L.kg_device_arch.restype = C.c_int
    L.kg_decode_workspace_bytes.restype = C.c_size_t
    L.kg_decode_workspace_bytes.argtypes = [C.POINTER(DecodeConfig), C.POINTER(DecodeScale)]
    L.kg_decode.restype = C.c_int
    L.kg_decode.argtypes = [C.POINTER(DecodeConfig), C.POINTER(DecodeScale), C.POINTER(DecodeOutputs), C.c_void_p,
                            C.c_size_t, C.c_void_p, C.POINTER(C.c_int)]
    L.kg_decode_host.restype = C.c_int
    L.kg_decode_host.argtypes = [C.POINTER(DecodeConfig), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.kg_skeletons_to_boxes_host.restype = C.c_int
    L.kg_skeletons_to_boxes_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_int)]
    L.kg_nms_host.restype = C.c_int
    L.kg_nms_host.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int)]
    L.kg_timing_enable.restype = C.c_int
    L.kg_timing_enable.argtypes = [C.c_int]
    L.kg_timing_collect.restype = C.c_int
    L.kg_timing_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    _bind_net(L)
    _lib = L
    return L


def _bind_net(L):
    """Signatures of the network entry points (present once the conv path is built in)."""
    vp, ci, cd, cs = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.kg_net_create.restype = ci
    L.kg_net_create.argtypes = [C.POINTER(vp), C.POINTER(ci)]
    L.kg_net_destroy.restype = None
    L.kg_net_destroy.argtypes = [vp]
    L.kg_net_set_conv.restype = ci
    L.kg_net_set_conv.argtypes = [vp, C.c_char_p, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp, cd]
    L.kg_net_finalize.restype = ci
    L.kg_net_finalize.argtypes = [vp]
    L.kg_net_workspace_bytes.restype = cs
    L.kg_net_workspace_bytes.argtypes = [vp, ci, ci, ci, ci]
    L.kg_net_forward_dec.restype = ci
    L.kg_net_forward_dec.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, vp, cs, vp, C.POINTER(ci)]
    L.kg_net_forward_dec_u8.restype = ci
    L.kg_net_forward_dec_u8.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, vp, cs, vp, C.POINTER(ci)]
    L.kg_net_import_feats.restype = ci
    L.kg_net_import_feats.argtypes = [vp, vp, ci, ci, ci, ci, vp, cs, vp]
    L.kg_net_seg_prepare.restype = ci
    L.kg_net_seg_prepare.argtypes = [vp, ci, ci, ci, vp, vp, C.POINTER(cs), C.POINTER(C.c_longlong), C.POINTER(ci), vp, vp, vp, vp]
    L.kg_net_forward_seg.restype = ci
    L.kg_net_forward_seg.argtypes = [vp, vp, vp, cs, vp, vp, C.POINTER(ci)]
    L.kg_conv2d_nchw.restype = ci
    L.kg_conv2d_nchw.argtypes = [vp, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, vp, vp]
    L.kg_heads_l2_nchw.restype = ci
    L.kg_heads_l2_nchw.argtypes = [vp, ci, ci, ci, ci, vp, vp, vp, vp]
    L.kg_net_plan_info.restype = ci
    L.kg_net_plan_info.argtypes = [vp, vp, ci]
    L.kg_preprocess_u8.restype = ci
    L.kg_preprocess_u8.argtypes = [vp, ci, ci, ci, vp, vp]
    L.kg_paste_masks.restype = ci
    L.kg_paste_masks.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, C.c_float, vp, vp, vp]
    L.kg_encode_ground_truth.restype = ci
    L.kg_encode_ground_truth.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    L.kg_detection_loss.restype = ci
    L.kg_detection_loss.argtypes = [vp, vp, vp, vp, ci, ci, ci, C.c_float, vp, vp, vp]
    L.kg_seg_loss_pairs.restype = ci
    L.kg_seg_loss_pairs.argtypes = [vp, vp, ci, vp, ci, ci, vp, vp]
    L.kg_detection_loss_backward.restype = ci
    L.kg_detection_loss_backward.argtypes = [vp, vp, vp, vp, ci, ci, ci, C.c_float, vp, vp, vp, vp, vp, vp]
    L.kg_seg_loss_pairs_backward.restype = ci
    L.kg_seg_loss_pairs_backward.argtypes = [vp, vp, ci, vp, ci, ci, vp, vp, vp]
    L.kg_adam_step.restype = ci
    L.kg_adam_step.argtypes = [vp, vp, ci, C.c_double, C.c_double, C.c_double, C.c_double, ci, vp]
    L.kg_debug_place_by_liveness.restype = ci
    L.kg_debug_place_by_liveness.argtypes = [ci, ci, vp, vp, vp, vp, vp]
    L.kg_tc_available.restype = ci
    L.kg_tc_status.restype = C.c_char_p


def check(code):
    if code != 0:
        raise KgError(code, lib().kg_last_error().decode(errors="replace"))


EXPORTS = ["kg_last_error", "kg_abi_version", "kg_device_arch", "kg_decode_workspace_bytes", "kg_decode",
           "kg_decode_host", "kg_skeletons_to_boxes_host", "kg_nms_host", "kg_timing_enable", "kg_timing_collect",
           "kg_net_create", "kg_net_destroy", "kg_net_set_conv", "kg_net_finalize", "kg_net_workspace_bytes", "kg_net_forward_dec", "kg_net_forward_dec_u8",
           "kg_net_import_feats", "kg_net_seg_prepare", "kg_net_forward_seg", "kg_conv2d_nchw", "kg_heads_l2_nchw", "kg_net_plan_info", "kg_tc_available", "kg_tc_status",
           "kg_preprocess_u8", "kg_paste_masks", "kg_encode_ground_truth", "kg_detection_loss", "kg_seg_loss_pairs",
           "kg_detection_loss_backward", "kg_seg_loss_pairs_backward", "kg_debug_place_by_liveness", "kg_adam_step"]


def timing_enable(on=True):
    check(lib().kg_timing_enable(int(on)))


def timing_collect(n_stages=64):
    """-> (ms[n_stages], launches[n_stages]) accumulated since the last collect (synchronises the device)."""
    import numpy as np
    ms = np.zeros(n_stages, np.float32); cnt = np.zeros(n_stages, np.int32)
    check(lib().kg_timing_collect(ms.ctypes.data, cnt.ctypes.data, n_stages))
    return ms, cnt
