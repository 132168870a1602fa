"""Generates tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

TEST INFRASTRUCTURE.  The reference publishes no golden vectors, so these files ARE the
pin: every array under `ref_*` was produced by /root/reference code (postprocessing.py, nms.py,
KGnet.py) on inputs that are stored next to it (decode) or regenerated from a seed by
oracle.kg_oracle.make_state_dict (the 74 M-parameter state dict is too large to commit).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import kg_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    if not hasattr(np, "int"):
        np.int = int
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import KGnet, postprocessing, nms  # noqa: E401
    return KGnet, postprocessing, nms


def _pack_skeletons(sks):
    return np.asarray(sks, np.float64).reshape(-1, 5, 3)


def decode_fixture(pp, nms, name, heads, store_inputs):
    out = {}
    refined = []
    for s, (kp, short, mid) in enumerate(heads):
        t = lambda a: torch.from_numpy(a[None])
        kph = np.ascontiguousarray(kp.transpose(1, 2, 0)); shh = np.ascontiguousarray(short.transpose(1, 2, 0))
        heat = pp.compute_heatmaps(kph, shh)
        from scipy.ndimage import gaussian_filter
        blur = np.stack([gaussian_filter(heat[:, :, i], sigma=2) for i in range(5)], -1)
        kps = pp.get_keypoints(blur, 0.004)
        sk = pp.get_skeletons_and_masks(t(kp), t(short), t(mid))
        if store_inputs:
            out[f"kp{s}"] = kp; out[f"short{s}"] = short; out[f"mid{s}"] = mid
            out[f"ref_heat{s}"] = heat.transpose(2, 0, 1).copy(); out[f"ref_blur{s}"] = blur.transpose(2, 0, 1).copy()
        out[f"ref_peak_id{s}"] = np.asarray([k["id"] for k in kps], np.int32)
        out[f"ref_peak_xy{s}"] = np.asarray([k["xy"] for k in kps], np.int32).reshape(-1, 2)
        out[f"ref_peak_conf{s}"] = np.asarray([k["conf"] for k in kps], np.float64)
        out[f"ref_skel{s}"] = _pack_skeletons(sk)
        r = pp.refine_skeleton(sk)
        out[f"ref_refined{s}"] = _pack_skeletons(r)
        refined.append(r)
    boxes = pp.gather_skeleton(*copy.deepcopy(refined))
    out["ref_boxes"] = np.asarray(boxes, np.float64).reshape(-1, 5)
    det = nms.non_maximum_suppression_numpy(boxes, 0.5)
    out["ref_dets"] = np.zeros((0, 5)) if det is None else det
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith("ref_") and "heat" not in k and "blur" not in k})


def main():
    KGnet, pp, nms = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    # 1. decode, inputs stored: 64x64 planted scene, 3 cells
    heads, boxes = O.planted_scene(7, 64, 64, 3, side=(24, 40), gap=6)
    decode_fixture(pp, nms, "decode_64_seed7.npz", heads, True)
    # 2. decode, inputs regenerated from the seed (oracle.planted_scene): 256x256, 20 cells
    heads, boxes = O.planted_scene(11, 256, 256, 20, side=(24, 80))
    decode_fixture(pp, nms, "decode_256_seed11.npz", heads, False)
    # 3. network: seeded calibrated weights, 1x3x64x64 input -> all head maps + feats
    sd = O.make_state_dict(seed=0)
    model = KGnet.resnet50(pretrained=False).eval()
    model.load_state_dict(sd, strict=True)
    torch.manual_seed(0)
    x = torch.rand(2, 3, 64, 64) - 0.5
    with torch.no_grad():
        ref = model.forward_dec(x)
        bx = [np.array([[4., 6., 40., 50., 0.9], [10., 10., 20., 22., 0.5], [0., 0., 2., 2., 0.1], [30., 30., 31., 31., .05]]),
              np.array([[0., 0., 63., 63., 0.7], [20., 5., 58., 30., 0.6]])]
        seg = model.forward_seg(ref[4], bx)
    out = {"x": x.numpy()}
    for s in range(4):
        for nme, a in zip(("kp", "short", "mid"), ref[s]):
            out[f"ref_{nme}{s}"] = a.numpy()
    for l, f in enumerate(ref[4]):
        out[f"ref_c{l}"] = f.numpy().astype(np.float16)     # features kept at half precision to bound the file size
    for i in range(2):
        out[f"boxes{i}"] = bx[i]
        for j, (p, d) in enumerate(zip(seg[0][i], seg[1][i])):
            out[f"ref_mask{i}_{j}"] = p.numpy(); out[f"ref_det{i}_{j}"] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "forward_64_seed0.npz"), **out)
    print("forward_64_seed0.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
