"""Parity at the BENCHMARKED configuration (VERDICT r1 "next" #1): the kernels take other code paths at 512x512 than on the
64x64 fixtures (strip mode, CTA pairs, shift-kernel row mode, TMA-store staging), so this file checks

  (a) forward_dec `fast` / `exact` against the oracle at 1x3x512x512 and on a 2-image 512x512 batch: kp <= 1e-3 (north_star),
  (b) the free-running pipeline (InstanceHeat.detect_batch on the network's OWN head maps) against the oracle's
      forward_dec -> decode -> NMS: bit-identical integer results when the decode is fed the same head maps, and identity of
      every well-separated peak when the head maps differ by the network's fp16 error (margin test, see below),
  (c) BASELINE config 4: 1024x1024, ~500 planted cells per image -- decode integers bit-exact,
  (d) the host-buffer entry point kg_decode_host and the reference-style test_inference / post_processing flow.
"""
import ctypes as C
import types

import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu

_cache = {}


def _report(name, obj):
    """Measured parity numbers go to gpurun_out/ (when it exists) so that they can be quoted in DESIGN.md."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, "data": obj}, default=str) + "\n")


def _oracle_512():
    """Oracle forward_dec of two seeded 512x512 images (a few seconds of CPU per image), shared by the tests below."""
    if "o512" not in _cache:
        sd = O.make_state_dict(seed=0)
        torch.manual_seed(0)
        x = torch.rand(2, 3, 512, 512) - 0.5
        _cache["o512"] = (sd, x, O.forward_dec(sd, x))
    return _cache["o512"]


def _model(precision, sd):
    from kg_instance_segmentation_b200 import KGnet
    m = KGnet.resnet50(pretrained=False, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("precision,kp_tol,off_tol", [("exact", 1e-4, 2e-3), ("fast", 1e-3, 2e-2)])
def test_forward_dec_512_matches_oracle(precision, kp_tol, off_tol):
    """(a) 2-image 512x512 batch and the first image alone; offsets tolerance is relative to max(1, |ref|max)."""
    sd, x, ref = _oracle_512()
    m = _model(precision, sd)
    out = m.forward_dec(x.cuda())
    single = m.forward_dec(x[:1].cuda())
    worst, bad = {}, []
    for s in range(4):
        for k, name in enumerate(("kp", "short", "mid")):
            got, r = out[s][k].cpu(), ref[s][k]
            assert got.shape == r.shape
            err = float((got - r).abs().max())
            tol = kp_tol if k == 0 else off_tol * max(1.0, float(r.abs().max()))
            worst[(s, name)] = err
            if not err <= tol:
                bad.append((precision, s, name, err, tol))
            assert torch.equal(single[s][k], out[s][k][:1]), "batch element 0 differs from the same image run alone"
    print(f"[{precision}] max |err| at 512x512:", {k: f"{v:.2e}" for k, v in worst.items()})
    _report(f"forward_dec_512[{precision}]", {f"{k[1]}{k[0]}": v for k, v in worst.items()})
    assert not bad, bad
    for l in range(5):
        r = ref[4][l]
        err = float((out[4][l].cpu() - r).abs().max())
        assert err <= 2e-3 * max(1.0, float(r.abs().max())), (l, err)


def _oracle_decode(heads_np):
    """heads_np: per scale (kp, short, mid) CHW arrays of ONE image -> (dets, per-scale peaks, per-scale heat [H,W,5])."""
    sks, pks, heats = [], [], []
    for kp, sh, mid in heads_np:
        s, p, h = O.decode_scale(kp, sh, mid)
        sks.append(O.refine_skeleton(s)); pks.append(p); heats.append(h)
    return O.nms(O.gather_skeleton(*sks), 0.5), pks, heats


def _peak_keys(p, H, W):
    return p["id"].astype(np.int64) * H * W + p["y"].astype(np.int64) * W + p["x"]


def _margins(heat, peaks, thresh=O.PEAK_THRESH):
    """For every peak: min(h - max 4-neighbour, h - thresh): how far the heat map may move before the peak disappears."""
    H, W, _ = heat.shape
    m = []
    for i, y, x in zip(peaks["id"], peaks["y"], peaks["x"]):
        h = heat[y, x, i]
        nb = [heat[yy, xx, i] for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)) if 0 <= yy < H and 0 <= xx < W]
        m.append(min(h - max(nb), h - thresh))
    return np.asarray(m)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_free_running_pipeline_512(precision):
    """(b) detect_batch on the calibrated network's own outputs, no teacher forcing."""
    from kg_instance_segmentation_b200 import postprocessing
    from kg_instance_segmentation_b200.inference import InstanceHeat
    sd, x, ref = _oracle_512()
    eng = InstanceHeat(model=_model(precision, sd), device="cuda:0")
    dets, seg = eng.detect_batch(x.cuda(), with_masks=True)
    out = eng.model.forward_dec(x.cuda())
    report = []
    for n in range(2):
        ours_np = [tuple(t[n].cpu().numpy() for t in out[s]) for s in range(4)]
        # 1. decode parity on the network's OWN head maps: every integer identical, confidences to 1e-9
        o_dets, o_pks, _ = _oracle_decode(ours_np)
        if o_dets is None:
            assert dets[n] is None
        else:
            assert dets[n] is not None and dets[n].shape == o_dets.shape, (n, None if dets[n] is None else dets[n].shape, o_dets.shape)
            assert np.array_equal(dets[n][:, :4], o_dets[:, :4]), "box corners differ from the oracle decode of the same head maps"
            np.testing.assert_allclose(dets[n][:, 4], o_dets[:, 4], rtol=1e-9)
        dbg = postprocessing.decode_batched([[t[n:n + 1] for t in out[s]] for s in range(4)], debug=True)
        # 2. through the network: the oracle's forward_dec -> decode against ours.  The head maps differ by the fp16 operand
        # error, so a peak can only be required where its margin exceeds the measured heat-map difference.
        ref_np = [tuple(t[n].numpy() for t in ref[s]) for s in range(4)]
        r_dets, r_pks, r_heats = _oracle_decode(ref_np)
        for s in range(4):
            H, W = ref_np[s][0].shape[1:]
            cnt = int(dbg.peak_count[0, s])
            ours_keys = set(dbg.peak_key[0, s, :cnt].cpu().numpy().tolist())
            assert ours_keys == set(_peak_keys(o_pks[s], H, W).tolist()), "peak keys differ from the oracle decode of the same head maps"
            ref_keys = _peak_keys(r_pks[s], H, W)
            heat_ours = dbg.heat[s][0].cpu().numpy().transpose(1, 2, 0)
            eps = float(np.abs(heat_ours - r_heats[s]).max())
            marg = _margins(r_heats[s], r_pks[s])
            solid = marg > 2 * eps
            missing = [k for k, ok in zip(ref_keys.tolist(), solid) if ok and k not in ours_keys]
            assert not missing, f"scale {s}: {len(missing)} well-separated reference peaks missing (eps={eps:.2e})"
            extra = ours_keys - set(ref_keys.tolist())
            if extra:   # a peak of ours that the reference lacks must be a near-tie in the reference's heat map
                ex = np.asarray(sorted(extra))
                pe = dict(id=ex // (H * W), y=(ex % (H * W)) // W, x=ex % W)
                assert (_margins(r_heats[s], pe) > -2 * eps).all(), f"scale {s}: spurious peaks beyond the heat-map error"
            report.append((n, s, len(ref_keys), len(ours_keys), len(set(ref_keys.tolist()) ^ ours_keys), f"{eps:.1e}"))
        same = (r_dets is None and dets[n] is None) or (r_dets is not None and dets[n] is not None and r_dets.shape == dets[n].shape
                                                         and np.array_equal(r_dets[:, :4], dets[n][:, :4]))
        report.append((n, "boxes identical to the reference's own network+decode", same))
    print(f"[{precision}] (image, scale, ref peaks, our peaks, symmetric difference, max heat err):", report)
    _report(f"free_running_512[{precision}]", report)
    # Identity of the raw peak SETS is not attainable through any re-implementation of the network: the calibrated random
    # net emits ~2 700 peaks per 512x512 image at scale 0, a fraction of a percent of which are near-ties that flip under
    # the 1e-6 heat-map difference of a different fp32 summation order (measured r02a, `exact`: 16 of 2 717; `fast`, whose
    # head maps differ by up to 8e-4 and heat maps by 4e-6: 319 of 2 717 -- the random net's heat maps are noise-like plateaus).
    # What is asserted above is the strongest true statement (every peak whose margin exceeds the measured heat difference is
    # reproduced, nothing spurious appears); here the flip rate is bounded.
    for r in report:
        if len(r) == 6:
            assert r[4] <= max(4, (0.01 if precision == "exact" else 0.2) * r[2]), r


def test_cfg4_dense_1024_decode_bit_exact():
    """(c) BASELINE config 4: 1024x1024, ~500 cells per image (postprocessing.py:98-124 at ~2 500 peaks per image-scale)."""
    from kg_instance_segmentation_b200 import postprocessing
    scenes = [O.planted_scene(50 + i, 1024, 1024, 500, side=(16, 40), gap=6)[0] for i in range(2)]
    batch = [tuple(torch.from_numpy(np.stack([sc[s][k] for sc in scenes])).cuda() for k in range(3)) for s in range(4)]
    res = postprocessing.decode_batched(batch, max_peaks=4096, max_boxes=4096, debug=True)
    dets = res.detections()
    for n, sc in enumerate(scenes):
        ref, sks, pks = O.decode_image(sc)
        assert len(pks[0]["id"]) > 2000, "the scene is not dense"
        for s in range(4):
            H, W = sc[s][0].shape[1:]
            order = np.argsort(-pks[s]["conf"], kind="stable")
            cnt = int(res.peak_count[n, s])
            assert cnt == len(order)
            assert np.array_equal(res.peak_key[n, s, :cnt].cpu().numpy(), _peak_keys(pks[s], H, W)[order]), (n, s)
        assert dets[n] is not None and dets[n].shape == ref.shape and len(ref) > 400
        assert np.array_equal(dets[n][:, :4], ref[:, :4])
        np.testing.assert_allclose(dets[n][:, 4], ref[:, 4], rtol=1e-9)


def test_decode_capacity_grows_instead_of_raising():
    """a list overflow re-runs with doubled capacity (the reference's lists are unbounded)."""
    from kg_instance_segmentation_b200 import postprocessing
    heads, _ = O.planted_scene(9, 256, 256, 30, side=(20, 40), gap=4)
    batch = [tuple(torch.from_numpy(a[None]).cuda() for a in h) for h in heads]
    res = postprocessing.decode_batched(batch, max_peaks=64, max_boxes=64)     # 30 cells x 5 keypoints > 64 peaks
    ref, _, _ = O.decode_image(heads)
    got = res.detections()[0]
    assert got.shape == ref.shape and np.array_equal(got[:, :4], ref[:, :4])


def test_kg_decode_host_entry_point():
    """(d) include/kgnet_b200.h kg_decode_host: host head maps in, host detections out."""
    from kg_instance_segmentation_b200 import _cabi
    L = _cabi.lib()
    scenes = [O.planted_scene(21 + i, 128, 128, 6, side=(24, 50))[0] for i in range(2)]
    arr = [[np.ascontiguousarray(np.stack([sc[s][k] for sc in scenes])) for s in range(4)] for k in range(3)]
    cfg = _cabi.DecodeConfig(2, 4, 512, 512, 0.5, 0.004)
    vp = lambda arrs: (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
    Hs = (C.c_int * 4)(*[a.shape[2] for a in arr[0]]); Ws = (C.c_int * 4)(*[a.shape[3] for a in arr[0]])
    sc = (C.c_int * 4)(1, 2, 4, 8)
    dets = np.zeros((2, 512, 5), np.float64); cnt = np.zeros(2, np.int32)
    _cabi.check(L.kg_decode_host(C.byref(cfg), vp(arr[0]), vp(arr[1]), vp(arr[2]), Hs, Ws, sc, dets.ctypes.data, cnt.ctypes.data, None))
    for n, scn in enumerate(scenes):
        ref, _, _ = O.decode_image(scn)
        assert cnt[n] == len(ref) and np.array_equal(dets[n, :cnt[n], :4], ref[:, :4])
        np.testing.assert_allclose(dets[n, :cnt[n], 4], ref[:, 4], rtol=1e-9)


def _paste_reference_check(preds, args, iw, ih):
    from kg_instance_segmentation_b200.inference import InstanceHeat
    eng = InstanceHeat.__new__(InstanceHeat)
    got = InstanceHeat.post_processing(eng, args, preds, iw, ih)
    ref = O.post_processing(preds, args.input_h, args.input_w, iw, ih, args.seg_thresh)
    assert got[0].dtype == np.float32 and got[0].shape == ref[0].shape and got[1].shape == ref[1].shape
    np.testing.assert_array_equal(got[1], ref[1])
    return got, ref


def test_post_processing_matches_reference_restatement():
    """test.py:127-157 on device (csrc/paste.cu) against the cv2 restatement: identical where no resize happens; where cv2
    interpolates, masks may differ only at pixels whose interpolated value sits within 1e-5 of the threshold (OpenCV's IPP
    path differs from its own generic path in the last ulp)."""
    import cv2
    rs = np.random.RandomState(0)
    args = types.SimpleNamespace(input_h=96, input_w=128, nms_thresh=0.5, seg_thresh=0.5)
    patches, dets = [], []
    for (y1, x1, y2, x2) in ((3.2, 4.7, 40.4, 60.6), (10.0, 20.0, 30.0, 50.0), (50.5, 100.5, 95.9, 127.8), (0.0, 0.0, 95.0, 127.0)):
        r = [int(np.round(np.float32(v))) for v in (y1, x1, y2, x2)]
        h, w = min(r[2], 95) - max(r[0], 0), min(r[3], 127) - max(r[1], 0)
        patches.append(torch.from_numpy(rs.rand(h, w).astype(np.float32)).cuda())
        dets.append(torch.Tensor([y1, x1, y2, x2, 0.5 + 0.1 * len(dets)]))
    patches.append(torch.from_numpy(rs.rand(17, 23).astype(np.float32)).cuda())       # patch size != rounded box: first resize is real
    dets.append(torch.Tensor([8.0, 9.0, 41.0, 70.0, 0.3]))
    preds = [[patches[:3], patches[3:]], [dets[:3], dets[3:]]]
    # same image size: both cv2.resize calls of boxes 0-3 are copies -> bit-identical masks
    got, ref = _paste_reference_check(preds, args, 128, 96)
    assert np.array_equal(got[0][:4], ref[0][:4])
    for iw, ih in ((128, 96), (200, 150), (64, 48), (301, 97)):
        got, ref = _paste_reference_check(preds, args, iw, ih)
        # pixels where the two disagree must be threshold near-ties of the cv2 canvas
        for k in range(len(patches)):
            if np.array_equal(got[0][k], ref[0][k]):
                continue
            y1, x1, y2, x2, _ = np.asarray(dets[k], np.float32)
            y1 = max(0, int(np.round(y1))); x1 = max(0, int(np.round(x1)))
            y2 = min(int(np.round(y2)), args.input_h - 1); x2 = min(int(np.round(x2)), args.input_w - 1)
            canvas = np.zeros((args.input_h, args.input_w), np.float32)
            canvas[y1:y2, x1:x2] = cv2.resize(patches[k].cpu().numpy(), (x2 - x1, y2 - y1))
            canvas = cv2.resize(canvas, (iw, ih))
            bad = got[0][k] != ref[0][k]
            assert (np.abs(canvas[bad] - args.seg_thresh) <= 1e-5).all(), (k, iw, ih, int(bad.sum()))


def test_test_inference_flow_matches_oracle():
    """(d) the reference-style flow: InstanceHeat(...).test_inference(args, image) with an HWC uint8 image
    (test.py:88-125), exact precision, against the oracle's restatement."""
    from kg_instance_segmentation_b200.inference import InstanceHeat
    sd = O.make_state_dict(seed=0)
    rs = np.random.RandomState(3)
    image = rs.randint(0, 256, size=(150, 210, 3)).astype(np.uint8)
    args = types.SimpleNamespace(input_h=128, input_w=128, nms_thresh=0.5, seg_thresh=0.5)
    eng = InstanceHeat(model=_model("exact", sd), device="cuda:0")
    boxes = eng.test_inference(args, image, bbox_flag=True)
    ref_boxes = O.test_inference(sd, image, 128, 128, bbox_flag=True)
    # head maps differ by ~1e-5 between the two networks: require the same boxes when the oracle finds any
    if ref_boxes is None:
        assert boxes is None or len(boxes) <= 2
    else:
        assert boxes is not None and boxes.shape == ref_boxes.shape and np.array_equal(boxes[:, :4], ref_boxes[:, :4])
    full = eng.test_inference(args, image)
    ref_full = O.test_inference(sd, image, 128, 128)
    if ref_full is None:
        assert full is None or len(full[0]) <= 2
    else:
        assert full[0].shape == ref_full[0].shape and full[0].shape[1:] == (150, 210)
        assert np.mean(full[0] != ref_full[0]) < 1e-3


def test_preprocess_u8_matches_reference_arithmetic():
    """test.py:92: FloatTensor(HWC->CHW)/255 - 0.5, bit-exact."""
    from kg_instance_segmentation_b200.inference import preprocess_u8
    rs = np.random.RandomState(1)
    img = rs.randint(0, 256, size=(3, 40, 56, 3)).astype(np.uint8)
    got = preprocess_u8(torch.from_numpy(img).cuda()).cpu()
    ref = torch.FloatTensor(np.transpose(img, (0, 3, 1, 2)).copy()) / 255 - 0.5
    assert torch.equal(got, ref)


def test_weight_update_after_forward_is_seen():
    """lazily packed weights (stem image, shift-add slabs) must follow load_state_dict."""
    sd0 = O.make_state_dict(seed=0)
    sd1 = O.make_state_dict(seed=1)
    torch.manual_seed(2)
    x = (torch.rand(1, 3, 64, 128) - 0.5).cuda()
    boxes = [np.array([[4., 6., 40., 100., 0.9]])]
    for prec in ("fast", "exact"):
        m = _model(prec, sd0)
        first = m.forward_dec(x)
        m.forward_seg(first[4], boxes)
        m.load_state_dict(sd1, strict=True)
        again = m.forward_dec(x)
        seg_again = m.forward_seg(again[4], boxes)
        fresh_m = _model(prec, sd1)
        fresh = fresh_m.forward_dec(x)
        seg_fresh = fresh_m.forward_seg(fresh[4], boxes)
        for s in range(4):
            for a, b in zip(again[s], fresh[s]):
                assert torch.equal(a, b), (prec, s)
        for l in range(5):
            assert torch.equal(again[4][l], fresh[4][l])
        assert torch.equal(seg_again[0][0][0], seg_fresh[0][0][0])
        assert not torch.equal(first[0][0], again[0][0])


def test_stale_feature_list_is_not_mistaken_for_the_current_pass():
    """with export_feats=False a feature list of an earlier forward_dec must not select the workspace."""
    sd = O.make_state_dict(seed=0)
    m = _model("fast", sd)
    m.export_feats = False
    torch.manual_seed(4)
    xa, xb = (torch.rand(1, 3, 64, 64) - 0.5).cuda(), (torch.rand(1, 3, 64, 64) - 0.5).cuda()
    fa = m.forward_dec(xa)[4]
    m.forward_dec(xb)
    with pytest.raises(RuntimeError):
        m.forward_seg(fa, [np.array([[4., 6., 40., 50., 0.9]])])


@pytest.mark.parametrize("precision,kp_tol", [("fast", 1e-3), ("exact", 1e-4)])
def test_forward_dec_and_seg_ragged_size(precision, kp_tol):
    """320 x 448: widths that are neither a multiple of the 128-pixel strip nor below it (scale 0: three full strips + a 64-pixel
    tail; scale 1: 224 = 128 + 96; scale 2: 112 < 128), 20 x 28 at the deepest level -- the ragged-tile paths of every kernel family."""
    sd = O.make_state_dict(seed=0)
    torch.manual_seed(11)
    x = torch.rand(1, 3, 320, 448) - 0.5
    ref = O.forward_dec(sd, x)
    m = _model(precision, sd)
    out = m.forward_dec(x.cuda())
    for s in range(4):
        assert float((out[s][0].cpu() - ref[s][0]).abs().max()) <= kp_tol, s
        for k in (1, 2):
            scale = max(1.0, float(ref[s][k].abs().max()))
            assert float((out[s][k].cpu() - ref[s][k]).abs().max()) <= (2e-2 if precision == "fast" else 2e-3) * scale, (s, k)
    boxes = [np.array([[3., 5., 150., 200., 0.9], [100., 300., 318., 446., 0.8], [10., 10., 14., 13., 0.4], [0., 0., 319., 447., 0.3]])]
    rseg = O.forward_seg(sd, ref[4], boxes)
    seg = m.forward_seg(out[4], boxes)
    assert len(seg[0][0]) == len(rseg[0][0]) == 4
    tol = 2e-2 if precision == "fast" else 2e-3
    for a, b in zip(seg[0][0], rseg[0][0]):
        a = a.cpu()
        assert a.shape == b.shape
        sure = (b - 0.5).abs() > tol
        assert bool(((a >= 0.5) == (b >= 0.5))[sure].all())
        if precision == "exact":
            assert float((a - b).abs().max()) <= tol
