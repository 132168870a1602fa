"""Seeded synthetic inputs for tests and benchmarks: a random reference-format state dict with
calibrated heads, and planted keypoint scenes (teacher-forced decode load).  Data generators only: nothing here
implements the inference path."""
from __future__ import annotations

import math

import numpy as np

from .config import EDGES, KP_RADIUS, BOX_SCALES as SCALES

DIR_EDGES = EDGES + [e[::-1] for e in EDGES]


def _t():
    import torch
    import torch.nn.functional as F
    return torch, F


def make_state_dict(seed=0, blocks=(3, 4, 6), calibrate=0.02, bn_jitter=True):
    """Random reference-format state dict: Kaiming fan_out normal conv weights (KGnet.py:212-217);
    every `*_head_c*.2.weight` scaled by `calibrate` so kp logits are O(1).  With bn_jitter the BN
    affine/running stats are randomised mildly so that BN folding is actually exercised."""
    torch, _ = _t()
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, bias):
        std = math.sqrt(2.0 / (co * k * k))
        sd[name + ".weight"] = torch.randn(co, ci, k, k, generator=g) * std
        if bias:
            bound = 1.0 / math.sqrt(ci * k * k)
            sd[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def bn(name, c):
        if bn_jitter:
            sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
            sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)
            sd[name + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
            sd[name + ".running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(c); sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c); sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("conv1", 64, 3, 7, False); bn("bn1", 64)
    inpl = 64
    for li, (planes, nb) in enumerate(zip((64, 128, 256), blocks)):
        for b in range(nb):
            p = f"layer{li + 1}.{b}"
            conv(p + ".conv1", planes, inpl, 1, False); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3, False); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1, False); bn(p + ".bn3", planes * 4)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1, False); bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    conv("c0_conv.0", 64, 3, 3, True); conv("c0_conv.2", 64, 64, 3, True)
    for l, (ci, co, cc) in enumerate(((64, 64, 128), (256, 64, 128), (512, 256, 512), (1024, 512, 1024))):
        conv(f"skip_combine.{l}.up.0", co, ci, 3, True); conv(f"skip_combine.{l}.cat_conv.0", co, cc, 1, True)
    conv("seg_head.0", 64, 64, 3, True); conv("seg_head.2", 1, 64, 3, True)
    conv("c4_up_conv.0", 512, 1024, 3, True); conv("c3_up_conv.0", 256, 512, 3, True)
    conv("c2_up_conv.0", 64, 256, 3, True); conv("c1_up_conv.0", 64, 64, 3, True)
    conv("c3_cat_refine.0", 512, 1024, 1, True); conv("c2_cat_refine.0", 256, 512, 1, True)
    conv("c1_cat_refine.0", 64, 128, 1, True); conv("c0_cat_refine.0", 64, 128, 1, True)
    for s, c in zip((3, 2, 1, 0), (512, 256, 64, 64)):
        for name, co in (("kp_head", 5), ("short_offset_head", 10), ("mid_offset_head", 40)):
            conv(f"{name}_c{s}.0", c, c, 7, True)
            conv(f"{name}_c{s}.2", co, c, 7, True)
            sd[f"{name}_c{s}.2.weight"] *= calibrate
    return sd


def planted_scene(seed, H, W, n_cells, side=(24, 110), gap=12, noise=0.3):
    """Teacher-forced decode load: non-overlapping boxes encoded with the semantics of
    preprocessing.get_ground_truth (preprocessing.py:45-118) at 4 scales, kp amplitude ~U(0.6,1) per
    instance-keypoint, offsets + N(0, noise).  Vectorised restatement (disc masks of radius KP_RADIUS,
    nearest-instance assignment, short offsets = centre - pixel inside discs, mid offsets = target kp
    - pixel inside the source kp disc).  Returns heads = [(kp[5,h,w], short[10,h,w], mid[40,h,w])]*4 f32
    and the planted boxes [n,4] (y1,x1,y2,x2) at scale 0."""
    rs = np.random.RandomState(seed)
    boxes = []
    tries = 0
    while len(boxes) < n_cells and tries < 200000:
        tries += 1
        h = rs.randint(side[0], side[1] + 1); w = rs.randint(side[0], side[1] + 1)
        y1 = rs.randint(2, max(3, H - h - 2)); x1 = rs.randint(2, max(3, W - w - 2))
        b = (y1, x1, y1 + h, x1 + w)
        if b[2] >= H - 1 or b[3] >= W - 1:
            continue
        if all(b[0] - gap > o[2] or o[0] - gap > b[2] or b[1] - gap > o[3] or o[1] - gap > b[3] for o in boxes):
            boxes.append(b)
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    heads = []
    for sc in SCALES:
        h, w = H // sc, W // sc
        bs = boxes / sc
        keep = ((bs[:, 2] - bs[:, 0]) > 2 * KP_RADIUS + 1) & ((bs[:, 3] - bs[:, 1]) > 2 * KP_RADIUS + 1)  # dataset_base.py:72
        bs = bs[keep]
        n = len(bs)
        kp = np.zeros((5, h, w), np.float32); short = np.zeros((10, h, w), np.float32); mid = np.zeros((40, h, w), np.float32)
        if n:
            y1, x1, y2, x2 = bs.T
            pts = np.stack([np.stack([x1, y1], 1), np.stack([x2, y1], 1), np.stack([x1, y2], 1), np.stack([x2, y2], 1),
                            np.stack([(x1 + x2) / 2, (y1 + y2) / 2], 1)], 1)            # [n,5,(x,y)]
            pts = np.floor(pts)
            amp = rs.uniform(0.6, 1.0, size=(n, 5))
            owner = np.full((5, h, w), -1, np.int64)
            R = KP_RADIUS
            for k in range(5):
                # nearest-instance assignment inside the discs, stamped window by window (a disc of radius R lies inside
                # the (2R+1)^2 window around its centre); strict '<' keeps the FIRST instance on distance ties, like argmin
                dist = np.full((h, w), np.inf)
                for j in range(n):
                    cx, cy = int(pts[j, k, 0]), int(pts[j, k, 1])
                    xa, xb, ya, yb = max(cx - R, 0), min(cx + R + 1, w), max(cy - R, 0), min(cy + R + 1, h)
                    if xa >= xb or ya >= yb:
                        continue
                    yy, xx = np.meshgrid(np.arange(ya, yb), np.arange(xa, xb), indexing="ij")
                    d = np.sqrt((xx - pts[j, k, 0]) ** 2 + (yy - pts[j, k, 1]) ** 2)
                    upd = (d <= R) & (d < dist[ya:yb, xa:xb])
                    dist[ya:yb, xa:xb] = np.where(upd, d, dist[ya:yb, xa:xb])
                    owner[k, ya:yb, xa:xb] = np.where(upd, j, owner[k, ya:yb, xa:xb])
                inside = owner[k] >= 0
                jj = np.clip(owner[k], 0, n - 1)
                yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
                kp[k] = np.where(inside, amp[jj, k], 0.0)
                short[2 * k] = np.where(inside, pts[jj, k, 0] - xx, 0.0)
                short[2 * k + 1] = np.where(inside, pts[jj, k, 1] - yy, 0.0)
            yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            for m, (a, b) in enumerate(DIR_EDGES):
                inside = owner[a] >= 0
                jj = np.clip(owner[a], 0, n - 1)
                mid[2 * m] = np.where(inside, pts[jj, b, 0] - xx, 0.0)
                mid[2 * m + 1] = np.where(inside, pts[jj, b, 1] - yy, 0.0)
        if noise > 0:
            short += rs.normal(0, noise, short.shape).astype(np.float32)
            mid += rs.normal(0, noise, mid.shape).astype(np.float32)
            kp = np.clip(kp + np.abs(rs.normal(0, 0.01, kp.shape)).astype(np.float32), 0, 1).astype(np.float32)
        heads.append((kp, short, mid))
    return heads, boxes
