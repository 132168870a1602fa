"""Drop-in for the reference's seg_loss.py: `SEG_loss(height, width)(predictions, gt_masks, gt_boxes)`.  The IoU matching of
predicted and ground-truth boxes is list logic and stays on the host (seg_loss.py:48-61, float32 like the reference's torch
scalars); every matched pair's mask BCE -- crop of the ground-truth mask, nearest-neighbour resize to the patch, mean binary
cross entropy -- runs in ONE launch of `seg_loss_kernel` (csrc/loss.cu) instead of a cv2.resize + H2D + BCE kernel per pair.
With reference-style lists of patch tensors that require grad, the result is differentiable with respect to the patches
(`seg_loss_backward_kernel`, one launch); patches of this library's own forward_seg carry no graph (no network backward here)."""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi

_PAIR = np.dtype([("patch_off", np.int64), ("pitch", np.int32), ("h", np.int32), ("w", np.int32), ("gt_index", np.int32),
                  ("y1", np.int32), ("x1", np.int32), ("y2", np.int32), ("x2", np.int32)], align=True)
assert _PAIR.itemsize == 40


def _jaccard(a, b):
    """seg_loss.py:14-29 on float32 scalars."""
    f = np.float32
    area_a = (a[2] - a[0]) * (a[3] - a[1])
    area_b = (b[2] - b[0]) * (b[3] - b[1])
    ih = max(min(a[2], b[2]) - max(a[0], b[0]), f(0.))
    iw = max(min(a[3], b[3]) - max(a[1], b[1]), f(0.))
    inter = ih * iw
    union = area_a + area_b - inter
    return f(0.) if union <= 2 else np.divide(inter, union)


class _PairLoss(torch.autograd.Function):
    """per-pair mean BCE of windows of `buf` (fp32, flat) against the nearest-resized ground-truth crops"""

    @staticmethod
    def forward(ctx, buf, d_rec, n_pairs, d_gt, height, width):
        dev = buf.device
        per_pair = torch.empty(n_pairs, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().kg_seg_loss_pairs(buf.data_ptr(), d_rec.data_ptr(), n_pairs, d_gt.data_ptr(), height, width,
                                                      per_pair.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(buf, d_rec, d_gt)
        ctx.geom = (n_pairs, height, width)
        return per_pair

    @staticmethod
    def backward(ctx, grad_pairs):
        buf, d_rec, d_gt = ctx.saved_tensors
        n_pairs, height, width = ctx.geom
        dev = buf.device
        coeff = grad_pairs.detach().to(device=dev, dtype=torch.float32).contiguous()
        grad = torch.zeros_like(buf)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().kg_seg_loss_pairs_backward(buf.data_ptr(), d_rec.data_ptr(), n_pairs, d_gt.data_ptr(), height, width,
                                                               coeff.data_ptr(), grad.data_ptr(),
                                                               torch.cuda.current_stream(dev).cuda_stream))
        return grad, None, None, None, None, None


class SEG_loss(torch.nn.Module):
    def __init__(self, height, width):
        super().__init__()
        self.height, self.width = height, width

    def forward(self, predictions, gt_masks, gt_boxes):
        """predictions = [mask_patches, mask_dets] of forward_seg; gt_masks[i]: (n_i, H, W) arrays (or list of HxW arrays);
        gt_boxes[i]: (n_i, >=4) rows [y1,x1,y2,x2,...].  Returns the scalar fp32 CUDA loss, or None when nothing matched."""
        mask_patches, mask_dets = predictions
        pairs, owners, gt_list, gt_slot = [], [], [], {}
        patches = []
        for i in range(len(mask_patches)):
            for j, patch in enumerate(mask_patches[i]):
                pbox = np.asarray(mask_dets[i][j].detach().cpu() if isinstance(mask_dets[i][j], torch.Tensor) else mask_dets[i][j], np.float32)[:4]
                gb = np.asarray(gt_boxes[i], np.float32).reshape(len(gt_boxes[i]), -1)
                for k in range(gb.shape[0]):
                    if _jaccard(pbox, gb[k]) >= 0.5:                                           # (:58-59)
                        y1 = max(0, int(np.int32(np.round(pbox[0])))); x1 = max(0, int(np.int32(np.round(pbox[1]))))
                        y2 = min(int(np.int32(np.round(pbox[2]))), self.height - 1)
                        x2 = min(int(np.int32(np.round(pbox[3]))), self.width - 1)             # (:60-64)
                        if (i, k) not in gt_slot:
                            gt_slot[(i, k)] = len(gt_list)
                            gt_list.append(np.asarray(gt_masks[i][k], np.float32))
                        pairs.append((len(patches), patch, gt_slot[(i, k)], y1, x1, y2, x2))
                        owners.append(i)
                patches.append(patch)
        if not pairs:
            return None                                                                         # run_label False (:93-96)
        dev = pairs[0][1].device
        packed = getattr(predictions, "packed", None)
        rec = np.zeros(len(pairs), _PAIR)
        if packed is not None:       # patches are windows of one device buffer (KGnet.SegResult)
            buf, off, pitch, hw = packed.paste_geometry()
            for r, (pi, _, g, y1, x1, y2, x2) in zip(rec, pairs):
                r["patch_off"], r["pitch"], r["h"], r["w"] = off[pi], pitch[pi], hw[pi, 0], hw[pi, 1]
                r["gt_index"], r["y1"], r["x1"], r["y2"], r["x2"] = g, y1, x1, y2, x2
        else:
            uniq = sorted({pi for pi, *_ in pairs})
            flat = [patches[pi].to(torch.float32).contiguous().view(-1) for pi in uniq]      # differentiable: grads flow back through cat
            starts = dict(zip(uniq, np.concatenate([[0], np.cumsum([f.numel() for f in flat])[:-1]]).tolist()))
            buf = torch.cat(flat)
            for r, (pi, patch, g, y1, x1, y2, x2) in zip(rec, pairs):
                r["patch_off"], r["pitch"], r["h"], r["w"] = starts[pi], patch.shape[1], patch.shape[0], patch.shape[1]
                r["gt_index"], r["y1"], r["x1"], r["y2"], r["x2"] = g, y1, x1, y2, x2
        for g in gt_list:
            if g.shape != (self.height, self.width):
                raise ValueError(f"ground-truth mask of shape {g.shape}, expected {(self.height, self.width)}")
        d_gt = torch.from_numpy(np.ascontiguousarray(np.stack(gt_list))).to(dev)
        d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(dev)
        per_pair = _PairLoss.apply(buf, d_rec, len(pairs), d_gt, self.height, self.width)
        # loss_batch / num_obj per image, then / len(mask_patches) (:88-94): a weighted sum of the per-pair terms
        owners = np.asarray(owners)
        counts = np.bincount(owners, minlength=len(mask_patches)).astype(np.float32)
        w = torch.from_numpy((1.0 / counts[owners] / np.float32(len(mask_patches))).astype(np.float32)).to(dev)
        return (per_pair * w).sum()
