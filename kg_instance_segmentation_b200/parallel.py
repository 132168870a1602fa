"""Data-parallel plumbing: images are independent end to end (no cross-image op in forward_dec, decode, NMS or
forward_seg), so the batch is sharded across ranks in contiguous slices and the ONLY collective of the path is one
all-gather of the fixed-size padded final detection list (torch.distributed: NCCL over NVLink on GPUs, gloo in the
CPU tests).  The reference has no multi-GPU path (its DataParallel stub is never called, test.py:57-58)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank`; the first (global_batch % world) ranks get one
    extra image."""
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pad_detections(dets: List[Optional[np.ndarray]], kmax: int, device=None):
    """Per-image (M,5) float64 arrays (or None) -> padded tensor [B, kmax, 5] f64 and counts [B] i32."""
    B = len(dets)
    out = torch.zeros(B, kmax, 5, dtype=torch.float64)
    cnt = torch.zeros(B, dtype=torch.int32)
    for i, d in enumerate(dets):
        if d is None:
            continue
        if len(d) > kmax:
            raise ValueError(f"image {i} has {len(d)} detections > kmax={kmax}")
        out[i, :len(d)] = torch.from_numpy(np.ascontiguousarray(d))
        cnt[i] = len(d)
    if device is not None:
        out, cnt = out.to(device), cnt.to(device)
    return out, cnt


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor, local_max: int, group=None):
    """dets [B_local, K, 5] f64, counts [B_local] i32 (device tensors for NCCL, CPU tensors for gloo).
    `local_max` = the largest per-rank batch (shards may differ by one image): shorter shards are padded so that every
    rank contributes the same number of bytes.  Returns (dets [world*local_max, K, 5], counts [world*local_max])."""
    world = dist.get_world_size(group)
    B, K, _ = dets.shape
    if B < local_max:
        dets = torch.cat([dets, dets.new_zeros(local_max - B, K, 5)], 0)
        counts = torch.cat([counts, counts.new_full((local_max - B,), -1)], 0)     # -1 marks padding slots
    g_d = dets.new_empty(world * local_max, K, 5)
    g_c = counts.new_empty(world * local_max)
    dist.all_gather_into_tensor(g_d, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(g_c, counts.contiguous(), group=group)
    return g_d, g_c


def trim_gathered(g_dets: torch.Tensor, g_counts: torch.Tensor) -> List[Optional[np.ndarray]]:
    """Gathered padded buffers -> ragged per-image list ordered by GLOBAL image index (padding slots dropped)."""
    d = g_dets.cpu().numpy()
    c = g_counts.cpu().numpy()
    out = []
    for i in range(len(c)):
        if c[i] < 0:
            continue
        out.append(d[i, :c[i]].copy() if c[i] > 0 else None)
    return out
