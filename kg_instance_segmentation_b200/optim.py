"""Drop-in for the optimizer of the reference's training loop (train.py:71: `torch.optim.Adam(params, lr=args.lr)`; :154
`optimizer.step()`): the update of ALL parameter tensors in ONE kernel launch (`adam_step_kernel`, csrc/optim.cu) instead of a
handful of elementwise launches per tensor.  PyTorch's defaults and operation order (torch/optim/adam.py, `_single_tensor_adam`);
works with `torch.optim.lr_scheduler` (train.py:72) because it is a `torch.optim.Optimizer` with the usual param_groups."""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi

_TENSOR = np.dtype([("p", np.uint64), ("g", np.uint64), ("m", np.uint64), ("v", np.uint64), ("n", np.int64)])
_CHUNK = np.dtype([("tensor", np.int32), ("pad", np.int32), ("start", np.int64)])
CHUNK = 65536


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError(f"invalid Adam hyper-parameters lr={lr} betas={betas} eps={eps}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._pinned = None
        self._copied = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            # parameters that received a gradient, bucketed by their own step count (PyTorch keeps it per parameter: a parameter
            # without a gradient does not advance); in the usual case there is one bucket, hence one launch
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("kg_instance_segmentation_b200.optim.Adam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("sparse gradients are not supported")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                buckets.setdefault((st["step"], p.device), []).append(p)
            for (step, dev), ps in buckets.items():
                grads = [p.grad if p.grad.is_contiguous() and p.grad.dtype == torch.float32 else p.grad.float().contiguous() for p in ps]
                tens = np.zeros(len(ps), _TENSOR)
                chunks = []
                for i, (p, g) in enumerate(zip(ps, grads)):
                    st = self.state[p]
                    tens[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                    chunks.extend((i, 0, s) for s in range(0, p.numel(), CHUNK))
                ch = np.array(chunks, _CHUNK) if chunks else np.zeros(0, _CHUNK)
                blob = np.concatenate([tens.view(np.uint8), ch.view(np.uint8)])
                if self._pinned is None or self._pinned.numel() < blob.size:
                    self._pinned = torch.empty(max(blob.size, 1 << 16), dtype=torch.uint8).pin_memory()
                with torch.cuda.device(dev):
                    stream = torch.cuda.current_stream(dev)
                    if self._copied is not None:
                        self._copied.synchronize()               # the previous table has left the pinned staging buffer
                    self._pinned[:blob.size].copy_(torch.from_numpy(blob))
                    d_blob = self._pinned[:blob.size].to(dev, non_blocking=True)
                    self._copied = torch.cuda.Event()
                    self._copied.record(stream)
                    _cabi.check(_cabi.lib().kg_adam_step(d_blob.data_ptr(), d_blob.data_ptr() + tens.nbytes, len(ch), float(group["lr"]),
                                                         float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), int(step),
                                                         stream.cuda_stream))
                    d_blob.record_stream(stream)
                del grads
        return loss
