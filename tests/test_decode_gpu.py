"""GPU parity tests of the decode path (vote -> blur -> peaks -> grouping -> boxes -> NMS) against the
oracle and the committed golden vectors.  Bar: integer results (peak ids/coordinates, skeleton membership,
box corners, keep order) bit-exact; fp64 confidences within CONF_RTOL (the vote accumulation is 2^-44
fixed point, see csrc/decode.cu)."""
import os

import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
HEAT_ATOL = 1e-11
CONF_RTOL = 1e-9


def _pp():
    from kg_instance_segmentation_b200 import postprocessing
    return postprocessing


def _to_batch(scenes):
    """scenes: list (per image) of heads [(kp,short,mid)]*S -> per scale stacked cuda tensors."""
    S = len(scenes[0])
    return [tuple(torch.from_numpy(np.stack([sc[s][k] for sc in scenes])).cuda() for k in range(3)) for s in range(S)]


def _check_image(res, n, heads, nms_thresh=0.5):
    for s, (kp, short, mid) in enumerate(heads):
        sk, pk, blur = O.decode_scale(kp, short, mid)
        H, W = kp.shape[1:]
        if res.vote is not None:
            vote = O.vote_heatmaps(np.ascontiguousarray(kp.transpose(1, 2, 0)), np.ascontiguousarray(short.transpose(1, 2, 0)))
            np.testing.assert_allclose(res.vote[s][n].cpu().numpy(), vote.transpose(2, 0, 1), rtol=0, atol=HEAT_ATOL)
            np.testing.assert_allclose(res.heat[s][n].cpu().numpy(), blur.transpose(2, 0, 1), rtol=0, atol=HEAT_ATOL)
        order = np.argsort(-pk["conf"], kind="stable")
        K = len(order)
        assert int(res.peak_count[n, s]) == K
        key = pk["id"].astype(np.int64) * H * W + pk["y"] * W + pk["x"]
        assert np.array_equal(res.peak_key[n, s, :K].cpu().numpy(), key[order])
        np.testing.assert_allclose(res.peak_conf[n, s, :K].cpu().numpy(), pk["conf"][order], rtol=CONF_RTOL)
        ns = int(res.skel_count[n, s])
        assert ns == len(sk)
        got = res.skeletons[n, s, :ns].cpu().numpy()
        ref = np.asarray(sk, np.float64).reshape(-1, 5, 3)
        assert np.array_equal(got[:, :, :2], ref[:, :, :2])
        np.testing.assert_allclose(got[:, :, 2], ref[:, :, 2], rtol=CONF_RTOL)
        keep = res.skel_keep[n, s, :ns].cpu().numpy().astype(bool)
        ref_keep = np.array([any(r is x for x in O.refine_skeleton(sk)) for r in sk], bool)
        assert np.array_equal(keep, ref_keep)
    det, sks, _ = O.decode_image(heads, nms_thresh)
    boxes = O.gather_skeleton(*sks).reshape(-1, 5)
    nb = int(res.box_count[n])
    assert nb == len(boxes)
    got = res.boxes[n, :nb].cpu().numpy()
    assert np.array_equal(got[:, :4], boxes[:, :4])
    np.testing.assert_allclose(got[:, 4], boxes[:, 4], rtol=CONF_RTOL)
    nd = int(res.det_count[n])
    assert nd == (0 if det is None else len(det))
    if nd:
        got = res.dets[n, :nd].cpu().numpy()
        assert np.array_equal(got[:, :4], det[:, :4])
        np.testing.assert_allclose(got[:, 4], det[:, 4], rtol=CONF_RTOL)


@pytest.mark.parametrize("seed,size,cells", [(1, 128, 6), (2, 192, 14), (3, 96, 4)])
def test_decode_stages_vs_oracle(seed, size, cells):
    heads, _ = O.planted_scene(seed, size, size, cells, side=(24, 60))
    res = _pp().decode_batched(_to_batch([heads]), debug=True, max_peaks=1024, max_boxes=1024)
    res.check()
    _check_image(res, 0, heads)


def test_decode_batch_of_ragged_scenes():
    """Images with different numbers of cells (including an empty one) in one batch."""
    scenes = [O.planted_scene(s, 128, 128, c, side=(24, 50))[0] for s, c in ((5, 7), (6, 1), (7, 3))]
    empty = [(np.zeros((5, 128 // k, 128 // k), np.float32), np.zeros((10, 128 // k, 128 // k), np.float32),
              np.zeros((40, 128 // k, 128 // k), np.float32)) for k in (1, 2, 4, 8)]
    scenes.insert(2, empty)
    res = _pp().decode_batched(_to_batch(scenes), debug=True, max_peaks=512, max_boxes=512)
    res.check()
    for n, heads in enumerate(scenes):
        _check_image(res, n, heads)
    dets = res.detections()
    assert dets[2] is None and all(d is not None for i, d in enumerate(dets) if i != 2)


def test_decode_golden_vectors():
    g = np.load(os.path.join(G, "decode_64_seed7.npz"))
    heads = [(g[f"kp{s}"], g[f"short{s}"], g[f"mid{s}"]) for s in range(4)]
    res = _pp().decode_batched(_to_batch([heads]), debug=True, max_peaks=256, max_boxes=256)
    res.check()
    for s in range(4):
        np.testing.assert_allclose(res.heat[s][0].cpu().numpy(), g[f"ref_blur{s}"], rtol=0, atol=HEAT_ATOL)
        ns = int(res.skel_count[0, s])
        got = res.skeletons[0, s, :ns].cpu().numpy()
        assert np.array_equal(got[:, :, :2], g[f"ref_skel{s}"][:, :, :2])
        np.testing.assert_allclose(got[:, :, 2], g[f"ref_skel{s}"][:, :, 2], rtol=CONF_RTOL)
    nb = int(res.box_count[0]); nd = int(res.det_count[0])
    assert np.array_equal(res.boxes[0, :nb, :4].cpu().numpy(), g["ref_boxes"][:, :4])
    assert np.array_equal(res.dets[0, :nd, :4].cpu().numpy(), g["ref_dets"][:, :4])
    np.testing.assert_allclose(res.dets[0, :nd, 4].cpu().numpy(), g["ref_dets"][:, 4], rtol=CONF_RTOL)

    g = np.load(os.path.join(G, "decode_256_seed11.npz"))
    heads, _ = O.planted_scene(11, 256, 256, 20, side=(24, 80))
    res = _pp().decode_batched(_to_batch([heads]), debug=True, max_peaks=1024, max_boxes=1024)
    res.check()
    for s in range(4):
        K = int(res.peak_count[0, s])
        order = np.argsort(-g[f"ref_peak_conf{s}"], kind="stable")
        H, W = heads[s][0].shape[1:]
        key = g[f"ref_peak_id{s}"].astype(np.int64) * H * W + g[f"ref_peak_xy{s}"][:, 1] * W + g[f"ref_peak_xy{s}"][:, 0]
        assert np.array_equal(res.peak_key[0, s, :K].cpu().numpy(), key[order])
    nd = int(res.det_count[0])
    assert np.array_equal(res.dets[0, :nd, :4].cpu().numpy(), g["ref_dets"][:, :4])


def test_reference_api_mirror_functions():
    pp = _pp()
    from kg_instance_segmentation_b200 import nms
    heads, _ = O.planted_scene(9, 160, 160, 8, side=(24, 60))
    mine, ref = [], []
    for kp, short, mid in heads:
        t = lambda a: torch.from_numpy(a[None])
        sk = pp.get_skeletons_and_masks(t(kp), t(short), t(mid))       # CPU tensors in, like the reference accepts
        rsk = O.decode_scale(kp, short, mid)[0]
        assert len(sk) == len(rsk) and all(np.array_equal(a[:, :2], b[:, :2]) for a, b in zip(sk, rsk))
        sk = pp.refine_skeleton(sk); rsk = O.refine_skeleton(rsk)
        assert len(sk) == len(rsk)
        mine.append(sk); ref.append(rsk)
    before = mine[1][0].copy() if len(mine[1]) else None
    boxes = pp.gather_skeleton(*mine); rboxes = O.gather_skeleton(*ref)
    if before is not None:
        assert np.array_equal(mine[1][0][:, :2], before[:, :2] * 2)    # the reference scales skeletons in place
    assert boxes.shape == rboxes.shape and np.array_equal(boxes[:, :4], rboxes[:, :4])
    np.testing.assert_allclose(boxes[:, 4], rboxes[:, 4], rtol=CONF_RTOL)
    out = nms.non_maximum_suppression_numpy(rboxes, 0.5)
    assert np.array_equal(out, O.nms(rboxes, 0.5))                    # same fp64 input -> bit-exact NMS
    assert nms.non_maximum_suppression_numpy(np.zeros((0,)), 0.5) is None
    assert pp.gather_skeleton([], [], [], []).shape == (0,)


def test_box_case_table_and_nms_edge_cases():
    pp = _pp()
    from kg_instance_segmentation_b200 import nms
    rs = np.random.RandomState(0)
    sks = []
    for mask in range(32):
        sk = np.zeros((5, 3))
        for k in range(5):
            if mask >> k & 1:
                sk[k] = (rs.randint(1, 60), rs.randint(0, 60), rs.uniform(0.01, 1))
        sks.append(sk)
    assert len(pp.refine_skeleton(sks)) == len(O.refine_skeleton(sks))
    for sc in (1, 2, 4, 8):
        mine = pp.skeleton_to_box([s.copy() for s in sks], sc)
        ref = O.skeleton_to_box(sks, sc)
        assert np.array_equal(np.asarray(mine), np.asarray(ref))
    # NMS: duplicates, zero-area boxes (0/0 -> NaN is dropped), containment, ties in conf
    b = np.array([[0, 0, 10, 10, .9], [0, 0, 10, 10, .8], [5, 5, 5, 5, .7], [5, 5, 5, 5, .95], [2, 2, 8, 8, .5],
                  [20, 20, 30, 30, .5], [21, 21, 31, 31, .5], [100, 100, 101, 150, .1]], np.float64)
    for thr in (0.3, 0.5, 0.9):
        assert np.array_equal(nms.non_maximum_suppression_numpy(b, thr), O.nms(b, thr))
    big = rs.uniform(0, 200, (1500, 5)); big[:, 2:4] = big[:, :2] + rs.uniform(5, 40, (1500, 2)); big[:, 4] = rs.uniform(0, 1, 1500)
    assert np.array_equal(nms.non_maximum_suppression_numpy(big, 0.5), O.nms(big, 0.5))


def test_nms_host_lists_beyond_the_shared_memory_kernel():
    """nms_host rejected more than 8192 boxes; longer lists now run the global-memory variant (same order, same
    arithmetic).  9 000 boxes on a wide canvas (few overlaps, so the oracle's greedy loop stays fast)."""
    from kg_instance_segmentation_b200 import nms
    rs = np.random.RandomState(5)
    big = rs.uniform(0, 4000, (9000, 5)); big[:, 2:4] = big[:, :2] + rs.uniform(5, 40, (9000, 2)); big[:, 4] = rs.uniform(0, 1, 9000)
    got = nms.non_maximum_suppression_numpy(big, 0.5)
    ref = O.nms(big, 0.5)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_overflow_is_reported_not_hidden():
    """A fixed-capacity Decoder reports the overflow in its status word; decode_batched re-runs with doubled capacity."""
    pp = _pp()
    heads, _ = O.planted_scene(11, 256, 256, 20, side=(24, 80))        # 70 peaks at scale 0 > cap of 64
    batch = [tuple(t.cuda() for t in h) for h in _to_batch([heads])]
    dec = pp.Decoder(1, [tuple(h[0].shape[2:]) for h in batch], max_peaks=64, max_boxes=64)
    res = dec(batch)
    assert res.overflow() & 1
    with pytest.raises(RuntimeError):
        res.check()
    grown = pp.decode_batched(batch, max_peaks=64, max_boxes=64)
    assert grown.overflow() == 0
    det, _, _ = O.decode_image(heads)
    assert np.array_equal(grown.detections()[0][:, :4], det[:, :4])


def test_full_size_batch_properties():
    """BASELINE config 2 sizes (bs 32, 512x512, 40 cells): run-to-run determinism (bit-exact, the vote is an
    integer accumulation), batch-permutation equivariance, and agreement of image 0 with the oracle."""
    pp = _pp()
    base = [O.planted_scene(100 + i, 512, 512, 40)[0] for i in range(4)]
    scenes = [base[i % 4] for i in range(32)]
    batch = _to_batch(scenes)
    r1 = pp.decode_batched(batch, max_peaks=4096, max_boxes=4096)
    r1.check()
    d1 = r1.dets.clone(); c1 = r1.det_count.clone()
    r2 = pp.decode_batched(batch, max_peaks=4096, max_boxes=4096)
    assert torch.equal(d1, r2.dets) and torch.equal(c1, r2.det_count)
    for i in range(4, 32):
        assert int(c1[i]) == int(c1[i % 4]) and torch.equal(d1[i, :int(c1[i])], d1[i % 4, :int(c1[i])])
    det, _, _ = O.decode_image(base[0])
    assert int(c1[0]) == len(det) and np.array_equal(d1[0, :len(det), :4].cpu().numpy(), det[:, :4])
    assert len(det) >= 30
