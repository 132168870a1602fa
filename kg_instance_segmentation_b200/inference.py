"""Inference orchestration: the reference's `InstanceHeat.test_inference` (test.py:88-125) for one image, and the
batched device pipeline `detect_batch` the benchmark times (forward_dec -> decode/group/NMS -> forward_seg)."""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from . import KGnet, postprocessing
from . import config as cfg


class InstanceHeat:
    def __init__(self, model=None, precision="fast", device="cuda:0"):
        self.device = torch.device(device)
        self.model = model if model is not None else KGnet.resnet50(pretrained=False, precision=precision)
        self.model.to(self.device).eval()
        self._decoders = {}
        self.last_launches = 0

    def load_weights(self, resume, dataset):
        """test.py:60-61."""
        self.model.load_state_dict(torch.load(os.path.join("weights_" + dataset, resume), map_location="cpu"))

    def _decoder(self, N, shapes, nms_thresh, max_peaks, max_boxes):
        key = (N, tuple(shapes), float(nms_thresh), max_peaks, max_boxes)
        d = self._decoders.get(key)
        if d is None:
            self._decoders.clear()
            d = self._decoders[key] = postprocessing.Decoder(N, shapes, nms_thresh=nms_thresh, max_peaks=max_peaks,
                                                             max_boxes=max_boxes, device=self.device)
        return d

    def detect_batch(self, x, nms_thresh=0.5, with_masks=True, head_override=None, max_peaks=4096, max_boxes=4096, packed=False):
        """x: [N,3,H,W] fp32 CUDA tensor in the reference's input convention (BGR/255 - 0.5, test.py:92).
        Returns (detections, seg): detections[i] = (M_i,5) float64 array or None (nms.py convention);
        seg = [mask_patches, mask_dets] of forward_seg, or the packed KGnet.SegResult when packed=True (None when
        with_masks is False).
        head_override: optional per-scale (kp, short, mid) CUDA tensors decoded INSTEAD of the network's own head
        outputs (teacher-forced decode load for benchmarking; the network still computes all of its heads)."""
        model = self.model
        keep = model.export_feats
        model.export_feats = False
        try:
            out = model.forward_dec(x)
        finally:
            model.export_feats = keep
        launches = model.last_launches
        heads = head_override if head_override is not None else [tuple(o) for o in out[:4]]
        N = x.shape[0]
        dec = self._decoder(N, [tuple(h[0].shape[2:]) for h in heads], nms_thresh, max_peaks, max_boxes)
        res = dec(heads)
        launches += res.n_launches
        dets = res.detections()                     # the one host sync of the pipeline: boxes are needed on the host
        self.last_result = res
        seg = None
        if with_masks:
            seg = model.forward_seg_packed(out[4], [d if d is not None else [] for d in dets])
            if not packed:
                seg = seg.as_lists()
            launches += model.last_launches
        self.last_launches = launches
        return dets, seg

    # ---- test.py:88-125 ---------------------------------------------------------------------------
    def test_inference(self, args, image, bbox_flag=False):
        import cv2
        height, width, c = image.shape
        img_input = cv2.resize(image, (args.input_w, args.input_h))
        img_input = torch.FloatTensor(np.transpose(img_input.copy(), (2, 0, 1))).unsqueeze(0) / 255 - 0.5
        img_input = img_input.to(self.device)
        dets, seg = self.detect_batch(img_input, nms_thresh=args.nms_thresh, with_masks=not bbox_flag)
        bboxes = dets[0]
        if bbox_flag:
            return bboxes
        if bboxes is None:
            return None
        return self.post_processing(args, seg, width, height)

    # ---- test.py:127-157 (host-side paste/resize with OpenCV, exactly as the reference does) ----
    def post_processing(self, args, predictions, image_w, image_h):
        import cv2
        if predictions is None:
            return predictions
        out_masks, out_dets = [], []
        mask_patches, mask_dets = predictions
        for mask_b_patches, mask_b_dets in zip(mask_patches, mask_dets):
            for mask_n_patch, mask_n_det in zip(mask_b_patches, mask_b_dets):
                mask_patch = mask_n_patch.data.cpu().numpy()
                y1, x1, y2, x2, conf = mask_n_det.data.cpu().numpy()
                y1 = np.maximum(0, np.int32(np.round(y1))); x1 = np.maximum(0, np.int32(np.round(x1)))
                y2 = np.minimum(np.int32(np.round(y2)), args.input_h - 1); x2 = np.minimum(np.int32(np.round(x2)), args.input_w - 1)
                mask = np.zeros((args.input_h, args.input_w), dtype=np.float32)
                mask[y1:y2, x1:x2] = cv2.resize(mask_patch, (x2 - x1, y2 - y1))
                mask = cv2.resize(mask, (image_w, image_h))
                mask = np.where(mask >= args.seg_thresh, 1, 0)
                out_masks.append(mask)
                out_dets.append([float(y1) / args.input_h * image_h, float(x1) / args.input_w * image_w,
                                 float(y2) / args.input_h * image_h, float(x2) / args.input_w * image_w, conf])
        return [np.asarray(out_masks, np.float32), np.asarray(out_dets, np.float32)]
