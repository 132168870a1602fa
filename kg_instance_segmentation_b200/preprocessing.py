"""Drop-in for the reference's preprocessing.py (ground-truth encoder of the training / validation input pipeline,
SURVEY.md 8f-3): same function name and return convention, computed by the CUDA kernel `gt_encode_kernel`
(csrc/gt_encode.cu) through the C-ABI entry point kg_encode_ground_truth.

`encode_ground_truth_batch` is the batched device entry (one launch for a whole batch of images at one scale, output
already in the [B, 55, H, W] layout DetectionLossAll consumes); `get_ground_truth` keeps the reference's per-image
NumPy contract."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _cabi
from . import config as cfg


def encode_ground_truth_batch(bboxes: Sequence[np.ndarray], height: int, width: int, device=None) -> torch.Tensor:
    """bboxes: per image an (n_i, 5, 2) array of keypoints (x, y) in the order tl, tr, bl, br, centre
    (dataset_base.masks_to_bboxes).  Returns the fp32 CUDA tensor [B, 55, height, width] =
    concat(kp heat, short offsets, mid offsets) of dataset_base.py:99-102 for every image."""
    if not torch.cuda.is_available():
        raise RuntimeError("kg_instance_segmentation_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    B = len(bboxes)
    counts = [0 if b is None or len(b) == 0 else len(b) for b in bboxes]
    off = np.zeros(B + 1, np.int32)
    off[1:] = np.cumsum(counts)
    rows = [np.asarray(b, np.float32).reshape(-1, cfg.NUM_KPS, 2) for b, c in zip(bboxes, counts) if c]
    flat = np.concatenate(rows, 0) if rows else np.zeros((1, cfg.NUM_KPS, 2), np.float32)
    d_boxes = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    gt = torch.empty(B, 5 + 2 * cfg.NUM_KPS + 4 * cfg.NUM_EDGES, height, width, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().kg_encode_ground_truth(d_boxes.data_ptr(), d_off.data_ptr(), B, int(height), int(width), gt.data_ptr(),
                                                       torch.cuda.current_stream(dev).cuda_stream))
    return gt


def get_ground_truth(bboxes, height, width, num_kps=cfg.NUM_KPS):
    """preprocessing.py:105-118: (kp_heats [num_kps,H,W], short_offsets [H,W,2*num_kps], mid_offsets [H,W,4*NUM_EDGES]),
    float64 NumPy arrays like the reference's."""
    if num_kps != cfg.NUM_KPS:
        raise ValueError(f"the keypoint graph has {cfg.NUM_KPS} keypoints (config.py)")
    gt = encode_ground_truth_batch([bboxes], height, width)[0].cpu().numpy().astype(np.float64)
    k = cfg.NUM_KPS
    return gt[:k], np.ascontiguousarray(gt[k:3 * k].transpose(1, 2, 0)), np.ascontiguousarray(gt[3 * k:].transpose(1, 2, 0))
