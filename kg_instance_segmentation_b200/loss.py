"""Drop-in for the reference's loss.py: `DetectionLossAll(kp_radius)(prediction, groundtruth)` computed by ONE pass of the CUDA
kernel `detection_loss_kernel` (csrc/loss.cu).  Forward only: this is the arithmetic of the reference's validation loop
(train.py:165-177); the returned tensor carries no autograd graph."""
from __future__ import annotations

import torch

from . import _cabi


class DetectionLossAll(torch.nn.Module):
    def __init__(self, kp_radius):
        super().__init__()
        self.kp_radius = kp_radius
        self.last_terms = None          # (kp BCE, short-offset, mid-offset) of the last call, device fp32 [3]

    def forward(self, prediction, groundtruth):
        """prediction = [pr_kp [N,5,H,W], pr_short [N,10,H,W], pr_mid [N,40,H,W]]; groundtruth [N,55,H,W] (loss.py:40-49).
        Returns the scalar fp32 CUDA tensor kp + short + 0.25 * mid."""
        pr = [t.detach().to(torch.float32).contiguous() for t in prediction]
        gt = groundtruth.detach().to(device=pr[0].device, dtype=torch.float32).contiguous()
        if not pr[0].is_cuda:
            raise RuntimeError("kg_instance_segmentation_b200 needs CUDA tensors (no CPU fallback)")
        N, _, H, W = pr[0].shape
        if tuple(gt.shape) != (N, 55, H, W) or tuple(pr[1].shape) != (N, 10, H, W) or tuple(pr[2].shape) != (N, 40, H, W):
            raise ValueError(f"shapes do not match: {[tuple(t.shape) for t in pr]} vs {tuple(gt.shape)}")
        dev = pr[0].device
        scratch = torch.empty(5, dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().kg_detection_loss(pr[0].data_ptr(), pr[1].data_ptr(), pr[2].data_ptr(), gt.data_ptr(), N, H, W,
                                                      float(self.kp_radius), scratch.data_ptr(), out.data_ptr(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
        self.last_terms = out[:3]
        return out[3]
