"""Drop-in for the reference's loss.py: `DetectionLossAll(kp_radius)(prediction, groundtruth)` computed by ONE pass of the CUDA
kernel `detection_loss_kernel` (csrc/loss.cu).  The returned scalar is differentiable with respect to the three prediction
tensors (`detection_loss_backward_kernel`, one more pass), so the module can replace the reference's in a `loss.backward()` training
loop (train.py:145-154) around any autograd network; the backward pass of this library's own network is not built."""
from __future__ import annotations

import torch

from . import _cabi


class _DetectionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pr_kp, pr_short, pr_mid, gt, kp_radius, holder):
        pr = [t.detach().to(torch.float32).contiguous() for t in (pr_kp, pr_short, pr_mid)]
        dev = pr[0].device
        N, _, H, W = pr[0].shape
        scratch = torch.empty(5, dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().kg_detection_loss(pr[0].data_ptr(), pr[1].data_ptr(), pr[2].data_ptr(), gt.data_ptr(), N, H, W,
                                                      float(kp_radius), scratch.data_ptr(), out.data_ptr(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(pr[0], pr[1], pr[2], gt, scratch)
        ctx.kp_radius = float(kp_radius)
        ctx.in_dtypes = (pr_kp.dtype, pr_short.dtype, pr_mid.dtype)
        holder.append(out[:3])
        return out[3].clone()

    @staticmethod
    def backward(ctx, grad_out):
        p0, p1, p2, gt, scratch = ctx.saved_tensors
        dev = p0.device
        N, _, H, W = p0.shape
        g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        grads = [torch.empty_like(p0), torch.empty_like(p1), torch.empty_like(p2)]
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().kg_detection_loss_backward(p0.data_ptr(), p1.data_ptr(), p2.data_ptr(), gt.data_ptr(), N, H, W,
                                                               ctx.kp_radius, scratch.data_ptr(), g.data_ptr(), grads[0].data_ptr(),
                                                               grads[1].data_ptr(), grads[2].data_ptr(),
                                                               torch.cuda.current_stream(dev).cuda_stream))
        need = ctx.needs_input_grad
        return tuple(gr.to(dt) if nd else None for gr, dt, nd in zip(grads, ctx.in_dtypes, need[:3])) + (None, None, None)


class DetectionLossAll(torch.nn.Module):
    def __init__(self, kp_radius):
        super().__init__()
        self.kp_radius = kp_radius
        self.last_terms = None          # (kp BCE, short-offset, mid-offset) of the last call, device fp32 [3]

    def forward(self, prediction, groundtruth):
        """prediction = [pr_kp [N,5,H,W], pr_short [N,10,H,W], pr_mid [N,40,H,W]]; groundtruth [N,55,H,W] (loss.py:40-49).
        Returns the scalar fp32 CUDA tensor kp + short + 0.25 * mid (with a grad_fn when a prediction requires grad)."""
        pr = list(prediction)
        if not pr[0].is_cuda:
            raise RuntimeError("kg_instance_segmentation_b200 needs CUDA tensors (no CPU fallback)")
        gt = groundtruth.detach().to(device=pr[0].device, dtype=torch.float32).contiguous()
        N, _, H, W = pr[0].shape
        if tuple(gt.shape) != (N, 55, H, W) or tuple(pr[1].shape) != (N, 10, H, W) or tuple(pr[2].shape) != (N, 40, H, W):
            raise ValueError(f"shapes do not match: {[tuple(t.shape) for t in pr]} vs {tuple(gt.shape)}")
        holder = []
        loss = _DetectionLoss.apply(pr[0], pr[1], pr[2], gt, self.kp_radius, holder)
        self.last_terms = holder[0]
        return loss
