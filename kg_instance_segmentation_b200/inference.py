"""Inference orchestration: the reference's `InstanceHeat.test_inference` / `post_processing` (test.py:88-157) and the
batched device pipeline `detect_batch` the benchmark times (preprocess -> forward_dec -> decode/group/NMS -> forward_seg
-> mask paste).  Everything between the host image and the final masks runs in the CUDA library."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np
import torch

from . import KGnet, postprocessing, _cabi
from . import config as cfg


def preprocess_u8(images: torch.Tensor) -> torch.Tensor:
    """uint8 NHWC CUDA batch (cv2 BGR order, already at the network size) -> fp32 NCHW `x / 255 - 0.5` (test.py:92)."""
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3 or not images.is_cuda:
        raise ValueError(f"expected a CUDA uint8 [N,H,W,3] batch, got {tuple(images.shape)} {images.dtype} {images.device}")
    images = images.contiguous()
    N, H, W, _ = images.shape
    x = torch.empty(N, 3, H, W, dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        _cabi.check(_cabi.lib().kg_preprocess_u8(images.data_ptr(), N, H, W, x.data_ptr(),
                                                 torch.cuda.current_stream(images.device).cuda_stream))
    return x


def paste_masks(patches: List[torch.Tensor], dets: np.ndarray, input_h, input_w, image_h, image_w, seg_thresh, packed=None):
    """Device version of the per-box loop of test.py:132-156.  patches: fp32 CUDA tensors (h_k x w_k), dets: (M,5) fp32 rows
    (y1,x1,y2,x2,conf).  Returns (masks uint8 CUDA [M,image_h,image_w], dets fp32 CUDA [M,5] in image coordinates)."""
    M = len(patches)
    dev = patches[0].device if M else torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(M, image_h, image_w, dtype=torch.uint8, device=dev)
    out_dets = torch.empty(M, 5, dtype=torch.float32, device=dev)
    if M == 0:
        return out, out_dets
    if packed is not None:        # patches are windows of ONE device buffer (SegResult): no gathering needed
        buf, off, pitch, hw = packed
    else:
        flat = [p.detach().to(torch.float32).contiguous().view(-1) for p in patches]
        buf = torch.cat(flat)
        sizes = np.asarray([f.numel() for f in flat], np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        hw = np.asarray([tuple(p.shape) for p in patches], np.int32).reshape(M, 2)
        pitch = hw[:, 1].copy()
    meta = np.concatenate([np.ascontiguousarray(off, np.int64).view(np.int32), np.ascontiguousarray(pitch, np.int32),
                           np.ascontiguousarray(hw, np.int32).reshape(-1),
                           np.ascontiguousarray(dets, np.float32).reshape(-1).view(np.int32)])
    d_meta = torch.from_numpy(meta).to(dev)                      # one small H2D for all the geometry
    p0 = d_meta.data_ptr()
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().kg_paste_masks(buf.data_ptr(), p0, p0 + 8 * M, p0 + 12 * M, p0 + 20 * M, M, int(input_h), int(input_w),
                                               int(image_h), int(image_w), float(seg_thresh), out.data_ptr(), out_dets.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
    return out, out_dets


class InstanceHeat:
    def __init__(self, model=None, precision="fast", device="cuda:0"):
        self.device = torch.device(device)
        self.model = model if model is not None else KGnet.resnet50(pretrained=True, precision=precision)   # test.py:53
        self.model.to(self.device).eval()
        self._decoders = {}      # slot -> {configuration: Decoder}
        self.last_launches = 0
        self.packed_k = 0        # > 0: the decode also writes the fixed-size per-image detection records (data-parallel all-gather)
        self._slots = None       # submit() / collect(): two engines (shared parameters) with their own workspaces
        self._n_submitted = self._n_collected = 0
        self.last_seg_model = self.model
        self.gap_events = None   # a list: detect_batch appends (decode finished, forward_seg starts) CUDA event pairs (diagnostics)

    def load_weights(self, resume, dataset):
        """test.py:60-61."""
        self.model.load_state_dict(torch.load(os.path.join("weights_" + dataset, resume), map_location="cpu"))

    def _decoder(self, N, shapes, nms_thresh, max_peaks, max_boxes, slot=0):
        key = (N, tuple(shapes), float(nms_thresh), max_peaks, max_boxes, self.packed_k)
        cache = self._decoders.setdefault(slot, {})
        d = cache.get(key)
        if d is None:
            cache.clear()
            d = cache[key] = postprocessing.Decoder(N, shapes, nms_thresh=nms_thresh, max_peaks=max_peaks,
                                                    max_boxes=max_boxes, device=self.device, packed_k=self.packed_k)
        return d

    # ---- two batches in flight ------------------------------------------------------------------------
    # detect_batch has one host round trip in the middle of the device work: the boxes must reach the host (the atlas of forward_seg
    # is planned there), so the GPU idles from the end of the decode until the first forward_seg launch (measured 1.2 ms per bs32
    # step).  submit() / collect() hide it: submit(i+1) enqueues forward_dec + decode of the NEXT batch before collect(i) waits for
    # the boxes of batch i, which by then landed in pinned memory long ago; forward_seg(i) queues up behind decode(i+1).  Two slots
    # (engine + workspaces + decode buffers each) alternate; the second engine shares the first one's parameter tensors.
    def _slot(self, k):
        if self._slots is not None and self._slots[0]["model"] is not self.model:      # the caller swapped engine.model
            if any(sl is not None and sl["pending"] is not None for sl in self._slots):
                raise RuntimeError("engine.model was replaced while a batch is in flight: collect() first")
            self._slots = None
        if self._slots is None:
            self._slots = [{"model": self.model, "pending": None}, None]
        if self._slots[k] is None:
            with torch.device("meta"):                      # no storage, no random initialisation: every tensor is replaced below
                twin = KGnet.resnet50(pretrained=False, precision=self.model.precision)
            src = dict(self.model.named_modules())
            for name, mod in twin.named_modules():           # the SAME Parameter / buffer objects: weight updates reach both engines
                for key in list(mod._parameters):
                    mod._parameters[key] = src[name]._parameters[key]
                for key in list(mod._buffers):
                    mod._buffers[key] = src[name]._buffers[key]
            twin._flat_tensors = None
            twin.eval()
            self._slots[k] = {"model": twin, "pending": None}
        slot = self._slots[k]
        slot["model"].precision = self.model.precision
        return slot

    def submit(self, x, nms_thresh=0.5, head_override=None, max_peaks=4096, max_boxes=4096, on_decoded=None):
        """Enqueue preprocess + forward_dec + decode of one batch (arguments as detect_batch) and return without waiting.  At most two
        batches may be in flight: collect() the older one before the third submit."""
        k = self._n_submitted & 1
        slot = self._slot(k)
        if slot["pending"] is not None:
            raise RuntimeError("two batches are already in flight: call collect() first")
        model = slot["model"]
        out, launches = self._forward_dec_any(model, x)
        heads = head_override if head_override is not None else [tuple(o) for o in out[:4]]
        N = x.shape[0]
        shapes = [tuple(h[0].shape[2:]) for h in heads]
        res = self._decoder(N, shapes, nms_thresh, max_peaks, max_boxes, slot=k)(heads)
        res.prefetch()                       # status / counts / boxes -> pinned memory, right behind the decode
        launches += res.n_launches
        if on_decoded is not None:
            on_decoded(res)
        slot["pending"] = dict(out=out, heads=heads, res=res, launches=launches, N=N, shapes=shapes, nms=nms_thresh,
                               caps=(max_peaks, max_boxes), on_decoded=on_decoded)
        self._n_submitted += 1

    def collect(self, with_masks=True, packed=False):
        """(detections, seg) of the OLDEST batch in flight, like detect_batch's return value."""
        k = self._n_collected & 1
        slot = self._slots[k] if self._slots is not None else None
        if slot is None or slot["pending"] is None:
            raise RuntimeError("collect() without a batch in flight")
        pend, model = slot["pending"], slot["model"]
        slot["pending"] = None
        self._n_collected += 1
        res, launches = pend["res"], pend["launches"]
        st = res.overflow() & 3
        if st:                               # a bounded device list overflowed: decode this batch again with doubled capacities
            mp, mb = pend["caps"]
            mp2 = min(2 * mp, postprocessing.MAX_CAPACITY) if st & 1 else mp
            mb2 = min(2 * mb, postprocessing.MAX_CAPACITY) if st & 2 else mb
            if (mp2, mb2) == (mp, mb):
                res.check()                  # raises: already at the largest capacity
            res = postprocessing.run_with_growth(lambda a, b: self._decoder(pend["N"], pend["shapes"], pend["nms"], a, b, slot=k),
                                                 pend["heads"], mp2, mb2)
            launches += res.n_launches
            if pend["on_decoded"] is not None:
                pend["on_decoded"](res)
        dets = res.detections()
        self.last_result = res
        seg = None
        if with_masks:
            seg = model.forward_seg_packed(pend["out"][4], [d if d is not None else [] for d in dets])
            if not packed:
                seg = seg.as_lists()
            launches += model.last_launches
        self.last_seg_model = model
        self.last_launches = launches
        return dets, seg

    def detect_pipelined(self, batches, **kw):
        """Generator over (detections, seg) of every batch of `batches`, two batches in flight."""
        collect_kw = {k: kw.pop(k) for k in ("with_masks", "packed") if k in kw}
        for x in batches:
            self.submit(x, **kw)
            if self._n_submitted - self._n_collected == 2:
                yield self.collect(**collect_kw)
        while self._n_submitted > self._n_collected:
            yield self.collect(**collect_kw)

    @staticmethod
    def _forward_dec_any(model, x):
        """forward_dec of an fp32 NCHW batch or of a uint8 NHWC batch (normalisation folded into the stem convs; through the separate
        normalisation kernel for precision 'reference').  Returns (outputs, launches); the feature maps stay inside the engine."""
        launches = 0
        keep = model.export_feats
        model.export_feats = False
        try:
            if x.dtype == torch.uint8 and model._precision_code() != 0:
                out = model.forward_dec_u8(x)
            else:
                if x.dtype == torch.uint8:
                    x = preprocess_u8(x)
                    launches += 1
                out = model.forward_dec(x)
        finally:
            model.export_feats = keep
        return out, launches + model.last_launches

    def detect_batch(self, x, nms_thresh=0.5, with_masks=True, head_override=None, max_peaks=4096, max_boxes=4096, packed=False,
                     on_decoded=None):
        """x: [N,3,H,W] fp32 CUDA tensor in the reference's input convention (BGR/255 - 0.5, test.py:92), or a uint8
        [N,H,W,3] CUDA batch (normalised on the device).
        Returns (detections, seg): detections[i] = (M_i,5) float64 array or None (nms.py convention);
        seg = [mask_patches, mask_dets] of forward_seg, or the packed KGnet.SegResult when packed=True (None when
        with_masks is False).
        head_override: optional per-scale (kp, short, mid) CUDA tensors decoded INSTEAD of the network's own head
        outputs (teacher-forced decode load for benchmarking; the network still computes all of its heads).
        on_decoded(result): called right after the decode has been ENQUEUED (before any host sync), e.g. to issue the
        data-parallel all-gather of result.packed on a side stream."""
        model = self.model
        out, launches = self._forward_dec_any(model, x)
        heads = head_override if head_override is not None else [tuple(o) for o in out[:4]]
        N = x.shape[0]
        shapes = [tuple(h[0].shape[2:]) for h in heads]
        res = postprocessing.run_with_growth(lambda mp, mb: self._decoder(N, shapes, nms_thresh, mp, mb), heads, max_peaks, max_boxes)
        launches += res.n_launches
        if on_decoded is not None:
            on_decoded(res)
        gap = None
        if self.gap_events is not None and with_masks:
            gap = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            gap[0].record()
        dets = res.detections()                     # the one host sync of the pipeline: boxes are needed on the host
        self.last_result = res
        self.last_seg_model = model
        seg = None
        if with_masks:
            model._seg_launch_event = gap[1] if gap is not None else None
            seg = model.forward_seg_packed(out[4], [d if d is not None else [] for d in dets])
            model._seg_launch_event = None
            if gap is not None and model.last_launches > 0:
                self.gap_events.append(gap)
            if not packed:
                seg = seg.as_lists()
            launches += model.last_launches
        self.last_launches = launches
        return dets, seg

    # ---- test.py:88-125 ---------------------------------------------------------------------------
    def test_inference(self, args, image, bbox_flag=False):
        """image: HWC uint8 (cv2.imread).  Returns [masks (M,h,w) f32, dets (M,5) f32], the boxes (bbox_flag) or None."""
        import cv2
        height, width = image.shape[:2]
        resized = cv2.resize(image, (args.input_w, args.input_h))                          # host, like the reference (:91)
        batch = torch.from_numpy(np.ascontiguousarray(resized)).unsqueeze(0).to(self.device)   # 3 B / pixel over PCIe
        dets, seg = self.detect_batch(batch, nms_thresh=args.nms_thresh, with_masks=not bbox_flag)
        if bbox_flag or dets[0] is None:
            return dets[0]
        return self.post_processing(args, seg, width, height)

    # ---- test.py:127-157 ---------------------------------------------------------------------------
    def post_processing(self, args, predictions, image_w, image_h):
        """predictions: [mask_patches, mask_dets] as returned by forward_seg.  One device launch (csrc/paste.cu) replaces
        the reference's per-box download + 2 x cv2.resize; the masks come back as ONE uint8 transfer."""
        if predictions is None:
            return predictions
        mask_patches, mask_dets = predictions
        patches = [p for per_image in mask_patches for p in per_image]
        rows = [np.asarray(d.detach().cpu() if isinstance(d, torch.Tensor) else d, np.float32) for per_image in mask_dets for d in per_image]
        if not patches:
            return [np.zeros((0,), np.float32), np.zeros((0,), np.float32)]        # np.asarray([]) twice (test.py:157)
        packed = getattr(predictions, "packed", None)
        geom = packed.paste_geometry() if packed is not None else None
        masks, dets = paste_masks(patches, np.stack(rows), args.input_h, args.input_w, image_h, image_w, args.seg_thresh, geom)
        return [masks.cpu().numpy().astype(np.float32), dets.cpu().numpy()]
