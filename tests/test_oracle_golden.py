"""The oracle must reproduce the committed golden vectors (made from the reference by oracle/gen_golden.py).
Runs everywhere (no reference checkout, no GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _check_decode(g, heads):
    refined = []
    for s, (kp, short, mid) in enumerate(heads):
        sk, pk, blur = O.decode_scale(kp, short, mid)
        if f"ref_blur{s}" in g.files:
            assert np.array_equal(blur.transpose(2, 0, 1), g[f"ref_blur{s}"])
        assert np.array_equal(pk["id"], g[f"ref_peak_id{s}"])
        assert np.array_equal(np.stack([pk["x"], pk["y"]], 1).reshape(-1, 2), g[f"ref_peak_xy{s}"])
        assert np.array_equal(pk["conf"], g[f"ref_peak_conf{s}"])
        assert np.array_equal(np.asarray(sk, np.float64).reshape(-1, 5, 3), g[f"ref_skel{s}"])
        r = O.refine_skeleton(sk)
        assert np.array_equal(np.asarray(r, np.float64).reshape(-1, 5, 3), g[f"ref_refined{s}"])
        refined.append(r)
    boxes = O.gather_skeleton(*refined)
    assert np.array_equal(boxes.reshape(-1, 5), g["ref_boxes"])
    det = O.nms(boxes, 0.5)
    assert np.array_equal(np.zeros((0, 5)) if det is None else det, g["ref_dets"])


def test_decode_golden_stored_inputs():
    g = np.load(os.path.join(G, "decode_64_seed7.npz"))
    _check_decode(g, [(g[f"kp{s}"], g[f"short{s}"], g[f"mid{s}"]) for s in range(4)])


def test_decode_golden_seeded_inputs():
    g = np.load(os.path.join(G, "decode_256_seed11.npz"))
    heads, _ = O.planted_scene(11, 256, 256, 20, side=(24, 80))
    _check_decode(g, heads)


def test_forward_golden():
    g = np.load(os.path.join(G, "forward_64_seed0.npz"))
    sd = O.make_state_dict(seed=0)
    out = O.forward_dec(sd, torch.from_numpy(g["x"]))
    for s in range(4):
        for nme, a in zip(("kp", "short", "mid"), out[s]):
            assert torch.equal(a, torch.from_numpy(g[f"ref_{nme}{s}"]))
    seg = O.forward_seg(sd, out[4], [g["boxes0"], g["boxes1"]])
    for i in range(2):
        for j, (p, d) in enumerate(zip(seg[0][i], seg[1][i])):
            assert torch.equal(p, torch.from_numpy(g[f"ref_mask{i}_{j}"]))
            assert torch.equal(d, torch.from_numpy(g[f"ref_det{i}_{j}"]))
        assert f"ref_mask{i}_{len(seg[0][i])}" not in g.files
