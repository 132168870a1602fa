"""Data-parallel plumbing: images are independent end to end (no cross-image op in forward_dec, decode, NMS or
forward_seg), so the batch is sharded across ranks in contiguous slices and the ONLY collective of the path is ONE
all-gather of the fixed-size per-image detection records (torch.distributed: NCCL over NVLink on GPUs, gloo in the
CPU tests).  The reference has no multi-GPU path (its DataParallel stub is never called, test.py:57-58).

Record layout (written on the device by nms_kernel, include/kgnet_b200.h `d_det_packed`): [B, K + 1, 5] f64 per rank;
row 0 = (detection count, rows stored, 0, 0, 0) -- count < 0 marks a padding slot of a short shard --, rows 1.. = the
detections [y1, x1, y2, x2, conf] in NMS keep order.  Counts travel inside the record: no second collective, no slicing."""
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


def pack_detections(dets: List[Optional[np.ndarray]], kmax: int, device=None) -> torch.Tensor:
    """Host-side builder of the record tensor (the device path gets it from the decode directly): per-image (M,5) float64
    arrays (or None) -> [B, kmax + 1, 5] f64."""
    B = len(dets)
    out = torch.zeros(B, kmax + 1, 5, dtype=torch.float64)
    for i, d in enumerate(dets):
        n = 0 if d is None else len(d)
        k = min(n, kmax)
        out[i, 0, 0], out[i, 0, 1] = n, k
        if k:
            out[i, 1:k + 1] = torch.from_numpy(np.ascontiguousarray(d[:k]))
    return out.to(device) if device is not None else out


def all_gather_records(records: torch.Tensor, local_max: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """records [B_local, K + 1, 5] f64 (device tensor for NCCL, CPU tensor for gloo).  `local_max` = the largest per-rank
    batch (shards may differ by one image): shorter shards are padded with count = -1 records so that every rank
    contributes the same number of bytes.  ONE collective.  Returns [world * local_max, K + 1, 5]."""
    world = dist.get_world_size(group)
    B = records.shape[0]
    if B < local_max:
        pad = records.new_zeros(local_max - B, *records.shape[1:])
        pad[:, 0, 0] = -1
        records = torch.cat([records, pad], 0)
    if out is None:
        out = records.new_empty(world * local_max, *records.shape[1:])
    dist.all_gather_into_tensor(out, records.contiguous(), group=group)
    return out


def unpack_records(gathered: torch.Tensor) -> List[Optional[np.ndarray]]:
    """Gathered records -> ragged per-image list ordered by GLOBAL image index (padding slots dropped); an image whose count
    exceeds the rows stored raises (the record was truncated: raise K)."""
    g = gathered.cpu().numpy()
    out = []
    for rec in g:
        n, k = int(rec[0, 0]), int(rec[0, 1])
        if n < 0:
            continue
        if k < n:
            raise ValueError(f"detection record truncated: {n} detections, {k} stored")
        out.append(rec[1:k + 1].copy() if k > 0 else None)
    return out
