"""CPU oracle for the KGnet inference hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy (decode, integer/fp64 work) and plain torch fp32 functional ops
(the conv net), what the reference computes on the path

    KGnet.forward_dec -> postprocessing.get_skeletons_and_masks -> refine_skeleton ->
    gather_skeleton -> nms.non_maximum_suppression_numpy -> KGnet.forward_seg

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import it; the product package (`kg_instance_segmentation_b200`) never does.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the *reference itself run in the authoring container* (`tests/test_oracle_vs_reference.py`,
skipped where /root/reference is absent) and against fixtures generated from the reference by
`oracle/gen_golden.py` and committed under `tests/golden/`.

Every function cites the reference file:line it follows (paths relative to the reference repo).
"""
from __future__ import annotations

import math

import numpy as np

# ---------------------------------------------------------------------------------------------
# config.py:2-19
EDGES = [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4)]
NUM_KPS = 5
KP_RADIUS = 5
PEAK_THRESH = 0.004          # postprocessing.py:145
SEED_SUPPRESS_RADIUS = 10.0  # postprocessing.py:100
MATCH_RADIUS = KP_RADIUS + 1  # postprocessing.py:114
GAUSS_SIGMA = 2.0            # postprocessing.py:144
GAUSS_RADIUS = 8             # scipy: int(truncate * sigma + 0.5), truncate = 4
SCALES = (1, 2, 4, 8)        # postprocessing.py:256-259

DIR_EDGES = EDGES + [e[::-1] for e in EDGES]  # postprocessing.py:89


def mid_index_table():
    """mid-offset edge index m for (seed id s -> target id t); targets visited in ascending t.

    postprocessing.py:68-78,91-96,105-109: BFS over the K5 skeleton graph from the seed reaches every
    other keypoint type at depth 1 in ascending order, so only edges (seed, t) are ever used.
    Returns int array [5, 4, 2] of (t, m).
    """
    tab = np.zeros((NUM_KPS, NUM_KPS - 1, 2), np.int32)
    for s in range(NUM_KPS):
        k = 0
        for t in range(NUM_KPS):
            if t == s:
                continue
            tab[s, k] = (t, DIR_EDGES.index((s, t)))
            k += 1
    return tab


def gaussian_weights():
    """scipy.ndimage._filters._gaussian_kernel1d(sigma=2, order=0, radius=8) (postprocessing.py:144)."""
    sigma2 = GAUSS_SIGMA * GAUSS_SIGMA
    x = np.arange(-GAUSS_RADIUS, GAUSS_RADIUS + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


# ---------------------------------------------------------------------------------------------
# Hough voting: postprocessing.py:8-53

def vote_heatmaps(kp_maps, short_offsets):
    """compute_heatmaps (postprocessing.py:39-53) + accumulate_votes (:16-37).

    kp_maps [H,W,5] f32, short_offsets [H,W,10] f32 (channel 2i = dx, 2i+1 = dy) -> [H,W,5] f64.
    coo_matrix(...).todense() (:36) adds duplicates sequentially in input order: all TL splats with
    pixels in row-major order, then TR, BL, BR; np.add.at is the same unbuffered sequential add.
    """
    H, W, K = kp_maps.shape
    ys_i, xs_i = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    out = np.zeros((H, W, K), np.float64)
    for i in range(K):
        xs = (xs_i + short_offsets[:, :, 2 * i]).astype(np.float64).reshape(-1)      # int64 + f32 -> f64 (:49)
        ys = (ys_i + short_offsets[:, :, 2 * i + 1]).astype(np.float64).reshape(-1)
        ps = kp_maps[:, :, i].astype(np.float64).reshape(-1)
        with np.errstate(invalid="ignore"):
            fy = np.floor(ys).astype(np.int32); fx = np.floor(xs).astype(np.int32)
            cy = np.ceil(ys).astype(np.int32); cx = np.ceil(xs).astype(np.int32)
        dx = xs - fx
        dy = ys - fy
        vals = np.concatenate([ps * (1. - dx) * (1. - dy), ps * dx * (1. - dy), ps * dy * (1. - dx), ps * dy * dx])
        I = np.concatenate([fy, fy, cy, cy])
        J = np.concatenate([fx, cx, fx, cx])
        good = (I >= 0) & (I < H) & (J >= 0) & (J < W)
        heat = np.zeros((H, W), np.float64)
        np.add.at(heat, (I[good], J[good]), vals[good])
        out[:, :, i] = heat / (np.pi * KP_RADIUS ** 2)
    return out


# ---------------------------------------------------------------------------------------------
# scipy.ndimage.gaussian_filter(sigma=2) restated: postprocessing.py:143-144

def _correlate1d_symmetric(a, w, axis):
    """scipy NI_Correlate1D symmetric branch, mode='reflect' (d c b a | a b c d | d c b a)."""
    r = (len(w) - 1) // 2
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    idx = np.arange(-r, n + r)
    # reflect without repeating pattern limits: period 2n
    idx = np.mod(idx, 2 * n)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    p = a[idx]
    tmp = p[r:r + n] * w[r]
    for j in range(-r, 0):
        tmp = tmp + (p[r + j:r + j + n] + p[r - j:r - j + n]) * w[j + r]
    return np.moveaxis(tmp, 0, axis)


def gaussian_blur(heat):
    """gaussian_filter per channel: axis 0 (y) pass then axis 1 (x) pass, fp64, no FMA."""
    w = gaussian_weights()
    out = np.empty_like(heat)
    for i in range(heat.shape[2]):
        t = _correlate1d_symmetric(heat[:, :, i], w, 0)
        out[:, :, i] = _correlate1d_symmetric(t, w, 1)
    return out


# ---------------------------------------------------------------------------------------------
# get_keypoints: postprocessing.py:56-64

def find_peaks(heat, peak_thresh=PEAK_THRESH):
    """Cross-footprint local maxima above threshold, in (id, y, x) generation order.

    Returns dict of arrays: id int32 [K], x int32 [K], y int32 [K], conf f64 [K].
    maximum_filter with the cross footprint pads with -inf-equivalent (reflect of a 1-px footprint
    never beats the centre), so out-of-image neighbours are ignored.
    """
    H, W, K = heat.shape
    ids, xs, ys, confs = [], [], [], []
    for i in range(K):
        h = heat[:, :, i]
        m = h.copy()
        m[1:, :] = np.maximum(m[1:, :], h[:-1, :])
        m[:-1, :] = np.maximum(m[:-1, :], h[1:, :])
        m[:, 1:] = np.maximum(m[:, 1:], h[:, :-1])
        m[:, :-1] = np.maximum(m[:, :-1], h[:, 1:])
        yy, xx = np.nonzero((m == h) & (h > peak_thresh))
        ids.append(np.full(len(yy), i, np.int32)); xs.append(xx.astype(np.int32)); ys.append(yy.astype(np.int32))
        confs.append(h[yy, xx])
    return dict(id=np.concatenate(ids), x=np.concatenate(xs), y=np.concatenate(ys), conf=np.concatenate(confs))


# ---------------------------------------------------------------------------------------------
# group_skeletons: postprocessing.py:80-126

def group_skeletons(peaks, mid_offsets):
    """Greedy keypoint-graph grouping.  peaks from find_peaks; mid_offsets [H,W,40] f32.

    Returns list of (5,3) f64 arrays [x, y, conf] (missing keypoints are all-zero rows).
    """
    K = len(peaks["id"])
    order = np.argsort(-peaks["conf"], kind="stable")       # list.sort(reverse=True) is stable (:87)
    pid = peaks["id"][order]; px = peaks["x"][order].astype(np.int64); py = peaks["y"][order].astype(np.int64)
    pc = peaks["conf"][order]
    alive = np.ones(K, bool)
    tab = mid_index_table()
    skeletons = []
    skel_xy = np.zeros((0, NUM_KPS, 2), np.float64)
    for i in range(K):
        if not alive[i]:
            continue
        alive[i] = False                                     # keypoints.pop(0) (:99)
        s = int(pid[i])
        if len(skeletons):
            d = np.sqrt((px[i] - skel_xy[:, s, 0]) ** 2 + (py[i] - skel_xy[:, s, 1]) ** 2)
            if np.any(d <= SEED_SUPPRESS_RADIUS):            # a missing kp sits at (0,0) (:100)
                continue
        sk = np.zeros((NUM_KPS, 3), np.float64)
        sk[s] = (px[i], py[i], pc[i])
        for t, m in tab[s]:
            off = mid_offsets[py[i], px[i], 2 * m:2 * m + 2]  # f32 (:110-112)
            prop_x = np.float64(px[i]) + np.float64(off[0])
            prop_y = np.float64(py[i]) + np.float64(off[1])
            cand = np.nonzero(alive & (pid == t))[0]
            if len(cand) == 0:
                continue
            ddx = prop_x - px[cand]; ddy = prop_y - py[cand]
            dist = np.sqrt(ddx * ddx + ddy * ddy)
            ok = dist <= MATCH_RADIUS
            if not ok.any():
                continue
            cand = cand[ok]; dist = dist[ok]
            j = cand[np.argmin(dist)]                         # stable sort by distance -> first minimum (:117)
            alive[j] = False
            sk[t] = (px[j], py[j], pc[j])
        skeletons.append(sk)
        skel_xy = np.concatenate([skel_xy, sk[None, :, :2]], 0)
    return skeletons


def decode_scale(kp, short, mid):
    """get_skeletons_and_masks (postprocessing.py:129-147) for ONE image: kp [5,H,W], short [10,H,W], mid [40,H,W]."""
    kp = np.ascontiguousarray(np.transpose(np.asarray(kp, np.float32), (1, 2, 0)))
    short = np.ascontiguousarray(np.transpose(np.asarray(short, np.float32), (1, 2, 0)))
    mid = np.ascontiguousarray(np.transpose(np.asarray(mid, np.float32), (1, 2, 0)))
    heat = gaussian_blur(vote_heatmaps(kp, short))
    peaks = find_peaks(heat)
    return group_skeletons(peaks, mid), peaks, heat


# ---------------------------------------------------------------------------------------------
# refine / boxes: postprocessing.py:150-261

def refine_skeleton(skeletons):
    """postprocessing.py:150-159: keep >=3 present keypoints or a diagonal pair; present = x > 0."""
    out = []
    for sk in skeletons:
        m = sk[:, 0] > 0.
        if m.sum() >= 3 or (m[0] and m[3]) or (m[1] and m[2]):
            out.append(sk)
    return out


def skeleton_to_box(skeletons, scale):
    """postprocessing.py:164-242 (does NOT mutate its input, unlike the reference)."""
    boxes = []
    for sk0 in skeletons:
        sk = sk0.copy()
        sk[:, :2] *= scale
        tl, tr, bl, br, cc = sk
        m = sk[:, 0] > 0.
        nc = int(m[:4].sum())
        conf = sk[m, 2].mean() if m.any() else 0.0
        if nc == 4:
            boxes.append([min(tl[1], tr[1]), min(tl[0], bl[0]), max(bl[1], br[1]), max(tr[0], br[0]), conf])
        elif nc == 3:
            y1 = min(tl[1], tr[1]) if (m[0] and m[1]) else max(tl[1], tr[1])
            x1 = min(tl[0], bl[0]) if (m[0] and m[2]) else max(tl[0], bl[0])
            boxes.append([y1, x1, max(bl[1], br[1]), max(tr[0], br[0]), conf])
        elif nc == 2:
            if m[0] and m[3]:
                boxes.append([tl[1], tl[0], br[1], br[0], conf])
            elif m[1] and m[2]:
                boxes.append([tr[1], bl[0], bl[1], tr[0], conf])
            elif m[0] and m[1] and m[4]:
                y1 = min(tl[1], tr[1])
                boxes.append([y1, tl[0], y1 + (cc[1] - y1) * 2, tr[0], conf])
            elif m[0] and m[2] and m[4]:
                x1 = min(tl[0], bl[0])
                boxes.append([tl[1], x1, bl[1], x1 + (cc[0] - x1) * 2, conf])
            elif m[1] and m[3] and m[4]:
                x2 = max(tr[0], br[0])
                boxes.append([tr[1], x2 - (x2 - cc[0]) * 2, br[1], x2, conf])
            elif m[2] and m[3] and m[4]:
                y2 = max(bl[1], br[1])
                boxes.append([y2 - (y2 - cc[1]) * 2, bl[0], y2, br[0], conf])
    return boxes


def gather_skeleton(s0, s1, s2, s3):
    """postprocessing.py:255-261."""
    b = []
    for sk, sc in zip((s0, s1, s2, s3), SCALES):
        b += skeleton_to_box(sk, sc)
    return np.asarray(b, np.float64).reshape(-1, 5) if len(b) else np.zeros((0,), np.float64)


# ---------------------------------------------------------------------------------------------
# nms.py:4-53

def nms(bboxes, nms_thresh=0.5):
    """Greedy NMS; returns rows in keep (descending conf) order, or None when input is empty.

    np.argsort(conf) (nms.py:16) is an unstable quicksort; ties between equal confs are resolved here
    by (conf, index) ascending, i.e. among equal confs the HIGHEST index is taken first.
    """
    if len(bboxes) == 0:
        return None
    y1, x1, y2, x2, conf = (bboxes[:, k] for k in range(5))
    area = (x2 - x1) * (y2 - y1)
    idx = np.lexsort((np.arange(len(conf)), conf))
    keep = []
    while len(idx) > 0:
        c = idx[-1]
        keep.append(c)
        if len(idx) == 1:
            break
        idx = idx[:-1]
        w = np.maximum(0., np.minimum(x2[idx], x2[c]) - np.maximum(x1[idx], x1[c]))
        h = np.maximum(0., np.minimum(y2[idx], y2[c]) - np.maximum(y1[idx], y1[c]))
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            iou = inter / ((area[idx] - inter) + area[c])
        idx = idx[iou <= nms_thresh]
    return bboxes[keep]


def decode_image(heads, nms_thresh=0.5):
    """test.py:105-116 for one image.  heads = [(kp,short,mid)]*4 of CHW f32 arrays.
    Returns (boxes (M,5) f64 or None, per-scale skeleton lists, per-scale peaks)."""
    sks, pks = [], []
    for kp, short, mid in heads:
        s, p, _ = decode_scale(kp, short, mid)
        sks.append(refine_skeleton(s)); pks.append(p)
    boxes = gather_skeleton(*sks)
    return nms(boxes, nms_thresh), sks, pks


# ---------------------------------------------------------------------------------------------
# KGnet.forward_dec restated with torch functional fp32 ops on a reference-format state dict.

def _t():
    import torch
    import torch.nn.functional as F
    return torch, F


def _bn(F, x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _bottleneck(F, x, sd, p, stride):
    """KGnet.py:64-99."""
    out = F.relu(_bn(F, F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F, F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F, F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(F, F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


def _convb(F, x, sd, p, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def forward_dec(sd, x, blocks=(3, 4, 6)):
    """KGnet.py:275-318.  sd: reference state dict (fp32 CPU tensors); x [N,3,H,W] fp32."""
    torch, F = _t()
    up = lambda a, ref: F.interpolate(a, ref.shape[2:], mode="bilinear", align_corners=False)
    with torch.no_grad():
        c0 = F.relu(_convb(F, F.relu(_convb(F, x, sd, "c0_conv.0", 1)), sd, "c0_conv.2", 1))
        c1 = F.relu(_bn(F, F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
        y = F.max_pool2d(c1, 3, 2, 1)
        feats = []
        for li, nb in enumerate(blocks):
            for b in range(nb):
                y = _bottleneck(F, y, sd, f"layer{li + 1}.{b}", 2 if (b == 0 and li > 0) else 1)
            feats.append(y)
        c2, c3, c4 = feats
        c4u = F.relu(_convb(F, up(c4, c3), sd, "c4_up_conv.0", 1))
        c3c = F.relu(_convb(F, torch.cat((c4u, c3), 1), sd, "c3_cat_refine.0", 0))
        c3u = F.relu(_convb(F, up(c3c, c2), sd, "c3_up_conv.0", 1))
        c2c = F.relu(_convb(F, torch.cat((c3u, c2), 1), sd, "c2_cat_refine.0", 0))
        c2u = F.relu(_convb(F, up(c2c, c1), sd, "c2_up_conv.0", 1))
        c1c = F.relu(_convb(F, torch.cat((c2u, c1), 1), sd, "c1_cat_refine.0", 0))
        c1u = F.relu(_convb(F, up(c1c, c0), sd, "c1_up_conv.0", 1))
        c0c = F.relu(_convb(F, torch.cat((c1u, c0), 1), sd, "c0_cat_refine.0", 0))
        outs = []
        for s, f in enumerate((c0c, c1c, c2c, c3c)):
            hd = lambda name: _convb(F, F.relu(_convb(F, f, sd, f"{name}_c{s}.0", 3)), sd, f"{name}_c{s}.2", 3)
            outs.append([torch.sigmoid(hd("kp_head")), hd("short_offset_head"), hd("mid_offset_head")])
    return outs[0], outs[1], outs[2], outs[3], [c0, c1, c2, c3, c4]


def get_patch_rect(box_norm, h, w):
    """KGnet.get_patches (KGnet.py:246-256) index arithmetic; box_norm are np.float32 scalars.
    Returns (y1,x1,y2,x2) int or None."""
    y1, x1, y2, x2 = box_norm
    y1 = np.maximum(0, np.int32(np.round(y1 * h)))
    x1 = np.maximum(0, np.int32(np.round(x1 * w)))
    y2 = np.minimum(np.int32(np.round(y2 * h)), h - 1)
    x2 = np.minimum(np.int32(np.round(x2 * w)), w - 1)
    if y2 < y1 or x2 < x1 or y2 - y1 < 2 or x2 - x1 < 2:
        return None
    return int(y1), int(x1), int(y2), int(x2)


def forward_seg(sd, feat_seg, bboxes):
    """KGnet.forward_seg (KGnet.py:321-350) restated.  Returns [mask_patches, mask_dets]."""
    torch, F = _t()
    mask_patches = [[] for _ in bboxes]
    mask_dets = [[] for _ in bboxes]
    with torch.no_grad():
        for i in range(len(bboxes)):
            if len(bboxes[i]) == 0:
                continue
            for box, score in zip(bboxes[i][:, :4], bboxes[i][:, 4]):
                y1, x1, y2, x2 = np.asarray(box, np.float32)
                h, w = feat_seg[0].shape[2:]
                patches = []
                for f in feat_seg:
                    r = get_patch_rect([y1 / float(h), x1 / float(w), y2 / float(h), x2 / float(w)],
                                       f.shape[2], f.shape[3])
                    if r is None:
                        break
                    patches.append(f[i:i + 1, :, r[0]:r[2], r[1]:r[3]])
                if not patches:
                    continue
                pre = patches[-1]
                for l in range(len(patches) - 2, -1, -1):       # mask_forward (KGnet.py:258-267)
                    p = f"skip_combine.{l}"
                    u = F.interpolate(pre, patches[l].shape[2:], mode="bilinear", align_corners=False)
                    u = F.relu(_convb(F, u, sd, p + ".up.0", 1))
                    pre = F.relu(_convb(F, torch.cat((patches[l], u), 1), sd, p + ".cat_conv.0", 0))
                if pre.shape[1] != 64:
                    # the reference would raise inside seg_head (only c0-level output has 64 channels);
                    # with >=1 patch the first is always c0, so pre has 64 channels here.
                    raise RuntimeError("seg_head expects 64 channels")
                xm = _convb(F, F.relu(_convb(F, pre, sd, "seg_head.0", 1)), sd, "seg_head.2", 1)
                mask_patches[i].append(torch.sigmoid(xm)[0, 0])
                mask_dets[i].append(torch.Tensor(np.append(box, score)))
    return [mask_patches, mask_dets]


# ---------------------------------------------------------------------------------------------
# InstanceHeat.post_processing / test_inference restated: test.py:88-157

def post_processing(predictions, input_h, input_w, image_w, image_h, seg_thresh):
    """test.py:127-157: every mask patch is resized to its rounded box, pasted into an input-sized canvas, the canvas is
    resized to the original image and thresholded.  Returns [masks (M,image_h,image_w) f32 in {0,1}, dets (M,5) f32]."""
    import cv2
    if predictions is None:
        return None
    masks, dets = [], []
    for patches_b, dets_b in zip(*predictions):
        for patch, det in zip(patches_b, dets_b):
            patch = np.asarray(patch.cpu() if hasattr(patch, "cpu") else patch, np.float32)
            y1, x1, y2, x2, conf = np.asarray(det.cpu() if hasattr(det, "cpu") else det, np.float32)
            y1 = np.maximum(0, np.int32(np.round(y1))); x1 = np.maximum(0, np.int32(np.round(x1)))             # :137-138
            y2 = np.minimum(np.int32(np.round(y2)), input_h - 1); x2 = np.minimum(np.int32(np.round(x2)), input_w - 1)
            canvas = np.zeros((input_h, input_w), np.float32)
            canvas[y1:y2, x1:x2] = cv2.resize(patch, (int(x2 - x1), int(y2 - y1)))                           # :143-146
            canvas = cv2.resize(canvas, (image_w, image_h))                                                  # :147
            masks.append(np.where(canvas >= seg_thresh, 1, 0))                                               # :148
            dets.append([float(y1) / input_h * image_h, float(x1) / input_w * image_w,
                         float(y2) / input_h * image_h, float(x2) / input_w * image_w, conf])
    return [np.asarray(masks, np.float32), np.asarray(dets, np.float32)]


def preprocess_image(image, input_h, input_w):
    """test.py:91-92: cv2.resize (bilinear) to the network size, HWC uint8 BGR -> [1,3,H,W] f32 in [-0.5, 0.5]."""
    import cv2
    torch, _ = _t()
    img = cv2.resize(image, (input_w, input_h))
    return torch.FloatTensor(np.transpose(img.copy(), (2, 0, 1))).unsqueeze(0) / 255 - 0.5


def test_inference(sd, image, input_h, input_w, nms_thresh=0.5, seg_thresh=0.5, bbox_flag=False):
    """test.py:88-125 on a reference-format state dict: image HWC uint8 -> [masks, dets] / boxes / None."""
    height, width, _ = image.shape
    out = forward_dec(sd, preprocess_image(image, input_h, input_w))
    heads = [tuple(t[0].numpy() for t in out[s]) for s in range(4)]
    boxes, _, _ = decode_image(heads, nms_thresh)
    if bbox_flag:
        return boxes
    if boxes is None:
        return None
    return post_processing(forward_seg(sd, out[4], [boxes]), input_h, input_w, width, height, seg_thresh)


# ---------------------------------------------------------------------------------------------
# Ground-truth encoder restated: preprocessing.py:45-118 (+ the channel concat of dataset_base.py:99-102)

def encode_ground_truth(bboxes, height, width):
    """bboxes [n,5,2] keypoints (x,y) -> gt [55,H,W] f32 = concat(kp heat [5], short offsets [10], mid offsets [40]).

    load_disc_masks (:62-77): nearest instance by fp64 Euclidean distance (np.argmin: first minimum), inside when <= KP_RADIUS.
    compute_short_offsets / copy_with_border_check (:12-60): every instance pastes its WHOLE clipped (2R+1)^2 window -- integer
    offsets int(centre) - pixel inside the radius-R circle, zeros in the corners; the mask line `temp_map[np.where(mask)==0,:] = 0.`
    is a no-op (a tuple compared with 0), so later instances overwrite earlier ones.
    compute_mid_offsets (:88-103): inside the disc of the edge's source keypoint: target keypoint of the same instance - pixel."""
    R = KP_RADIUS
    b = np.asarray(bboxes, np.float32).reshape(-1, NUM_KPS, 2)
    n = len(b)
    gt = np.zeros((5 + 2 * NUM_KPS + 4 * len(EDGES), height, width), np.float64)
    if n == 0:
        return gt.astype(np.float32)
    yy, xx = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    owner = np.full((NUM_KPS, height, width), -1, np.int64)
    for i in range(NUM_KPS):
        d = np.sqrt((b[:, i, 0].astype(np.float64)[:, None, None] - xx[None]) ** 2 + (b[:, i, 1].astype(np.float64)[:, None, None] - yy[None]) ** 2)
        j = d.argmin(0)
        inside = np.take_along_axis(d, j[None], 0)[0] <= R
        owner[i] = np.where(inside, j, -1)
        gt[i] = inside
        for k in range(n):                                   # short offsets: window paste in instance order
            cx, cy = int(b[k, i, 0]), int(b[k, i, 1])
            y1, y2 = max(cy - R, 0), min(cy + R, height - 1) + 1
            x1, x2 = max(cx - R, 0), min(cx + R, width - 1) + 1
            if y2 <= y1 or x2 <= x1:
                continue
            ox = cx - xx[y1:y2, x1:x2]; oy = cy - yy[y1:y2, x1:x2]
            circ = ox * ox + oy * oy <= R * R
            gt[5 + 2 * i, y1:y2, x1:x2] = ox * circ
            gt[5 + 2 * i + 1, y1:y2, x1:x2] = oy * circ
    for m, (a, t) in enumerate(DIR_EDGES):
        inside = owner[a] >= 0
        jj = np.clip(owner[a], 0, n - 1)
        gt[15 + 2 * m] = np.where(inside, b[jj, t, 0].astype(np.float64) - xx, 0.0)
        gt[15 + 2 * m + 1] = np.where(inside, b[jj, t, 1].astype(np.float64) - yy, 0.0)
    return gt.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# Losses restated (forward): loss.py:6-49, seg_loss.py:8-97

def detection_loss(prediction, groundtruth, kp_radius=KP_RADIUS):
    """DetectionLossAll.forward (loss.py:40-49) -> (total, kp, short, mid) torch scalars."""
    torch, F = _t()
    pr_kp, pr_short, pr_mid = prediction
    gt_kp, gt_short, gt_mid = groundtruth[:, :5], groundtruth[:, 5:15], groundtruth[:, 15:]
    kp = F.binary_cross_entropy(pr_kp, gt_kp)                                                 # :12-14
    m2 = gt_kp.repeat_interleave(2, dim=1)                                                   # :18-23
    short = (torch.abs(pr_short - gt_short) / kp_radius * m2).sum() / (m2.sum() + 1e-10)    # :17,24-25
    src = [e[0] for e in DIR_EDGES]
    m4 = gt_kp[:, src].repeat_interleave(2, dim=1)                                           # :30-35
    mid = (torch.abs(pr_mid - gt_mid) / kp_radius * m4).sum() / (m4.sum() + 1e-10)          # :29,36-37
    return kp + short + 0.25 * mid, kp, short, mid


def seg_loss(predictions, gt_masks, gt_boxes, height, width):
    """SEG_loss.forward (seg_loss.py:31-97) -> torch scalar or None."""
    import cv2
    torch, F = _t()
    f = np.float32

    def jaccard(a, b):                                                                        # :14-29
        area_a = (a[2] - a[0]) * (a[3] - a[1]); area_b = (b[2] - b[0]) * (b[3] - b[1])
        ih = max(min(a[2], b[2]) - max(a[0], b[0]), f(0.)); iw = max(min(a[3], b[3]) - max(a[1], b[1]), f(0.))
        inter = ih * iw
        union = area_a + area_b - inter
        return f(0.) if union <= 2 else np.divide(inter, union)

    mask_patches, mask_dets = predictions
    total, any_match = 0., False
    for i in range(len(mask_patches)):
        loss_batch, num_obj = 0., 0
        for j in range(len(mask_patches[i])):
            patch = mask_patches[i][j].cpu().to(torch.float32)          # (keeps the autograd graph of a patch that requires grad)
            pbox = np.asarray(mask_dets[i][j], np.float32)[:4]
            for k in range(len(gt_boxes[i])):
                if jaccard(pbox, np.asarray(gt_boxes[i][k], np.float32)) >= 0.5:
                    y1 = np.maximum(0, np.int32(np.round(pbox[0]))); x1 = np.maximum(0, np.int32(np.round(pbox[1])))
                    y2 = np.minimum(np.int32(np.round(pbox[2])), height - 1); x2 = np.minimum(np.int32(np.round(pbox[3])), width - 1)
                    crop = np.asarray(gt_masks[i][k], np.float32)[y1:y2, x1:x2]
                    h1, w1 = patch.shape
                    crop = cv2.resize(crop, (w1, h1), interpolation=cv2.INTER_NEAREST)        # :76
                    loss_batch = loss_batch + F.binary_cross_entropy(patch, torch.from_numpy(crop))
                    num_obj += 1
                    any_match = True
        if num_obj:
            total = total + loss_batch / num_obj
    return total / len(mask_patches) if any_match else None


# ---------------------------------------------------------------------------------------------
# Synthetic inputs live in the product package (they are data generators, not the algorithm);
# re-exported here because the tests and golden generator historically call them through the oracle.
from kg_instance_segmentation_b200.synthetic import make_state_dict, planted_scene  # noqa: E402,F401
