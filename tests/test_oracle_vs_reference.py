"""Pins oracle/kg_oracle.py against the UNMODIFIED reference run in the authoring container.

Skipped where /root/reference is absent (the GPU box); there the committed golden vectors
(tests/golden/, produced from the reference by oracle/gen_golden.py) take over.
"""
import copy

import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

pytestmark = pytest.mark.reference


def _ref_decode_scale(pp, kp, short, mid):
    t = lambda a: torch.from_numpy(a[None])
    return pp.get_skeletons_and_masks(t(kp), t(short), t(mid))


@pytest.mark.parametrize("seed,size,cells", [(1, 128, 6), (2, 192, 14)])
def test_decode_stages_bit_exact(reference, seed, size, cells):
    _, pp, nms = reference
    heads, _ = O.planted_scene(seed, size, size, cells, side=(24, 60))
    all_ref, all_ora = [], []
    for kp, short, mid in heads:
        kph = np.ascontiguousarray(kp.transpose(1, 2, 0)); shh = np.ascontiguousarray(short.transpose(1, 2, 0))
        ref_heat = pp.compute_heatmaps(kph, shh)
        ora_heat = O.vote_heatmaps(kph, shh)
        assert np.array_equal(ref_heat, ora_heat)
        from scipy.ndimage import gaussian_filter
        ref_blur = np.stack([gaussian_filter(ref_heat[:, :, i], sigma=2) for i in range(5)], -1)
        ora_blur = O.gaussian_blur(ora_heat)
        assert np.array_equal(ref_blur, ora_blur)
        ref_kps = pp.get_keypoints(ref_blur, 0.004)
        pk = O.find_peaks(ora_blur)
        assert [k["id"] for k in ref_kps] == pk["id"].tolist()
        assert [tuple(k["xy"]) for k in ref_kps] == list(zip(pk["x"].tolist(), pk["y"].tolist()))
        assert [k["conf"] for k in ref_kps] == pk["conf"].tolist()
        ref_sk = _ref_decode_scale(pp, kp, short, mid)
        ora_sk, _, _ = O.decode_scale(kp, short, mid)
        assert len(ref_sk) == len(ora_sk)
        for a, b in zip(ref_sk, ora_sk):
            assert np.array_equal(a, b)
        all_ref.append(pp.refine_skeleton(ref_sk)); all_ora.append(O.refine_skeleton(ora_sk))
    ref_boxes = pp.gather_skeleton(*copy.deepcopy(all_ref))
    ora_boxes = O.gather_skeleton(*all_ora)
    assert ref_boxes.shape == ora_boxes.shape and np.array_equal(ref_boxes, ora_boxes)
    assert len(ora_boxes) > 0
    r = nms.non_maximum_suppression_numpy(ref_boxes, 0.5); o = O.nms(ora_boxes, 0.5)
    assert np.array_equal(r, o)


def test_empty_and_edge_cases(reference):
    _, pp, nms = reference
    assert nms.non_maximum_suppression_numpy(np.zeros((0,)), 0.5) is None and O.nms(np.zeros((0,)), 0.5) is None
    z = [np.zeros((c, 32, 32), np.float32) for c in (5, 10, 40)]
    assert _ref_decode_scale(pp, *z) == [] and O.decode_scale(*z)[0] == []
    # single keypoint near the origin + a keypoint with x == 0 ("absent" by the x>0 rule)
    kp = np.zeros((5, 32, 32), np.float32); kp[0, 4, 0] = 1.0; kp[3, 20, 20] = 0.9; kp[1, 3, 3] = 0.8
    ref = _ref_decode_scale(pp, kp, z[1], z[2]); ora = O.decode_scale(kp, z[1], z[2])[0]
    assert len(ref) == len(ora) and all(np.array_equal(a, b) for a, b in zip(ref, ora))
    assert len(pp.refine_skeleton(ref)) == len(O.refine_skeleton(ora))


def test_box_case_table(reference):
    _, pp, _ = reference
    rs = np.random.RandomState(0)
    sks = []
    for mask in range(32):
        sk = np.zeros((5, 3))
        for k in range(5):
            if mask >> k & 1:
                sk[k] = (rs.randint(1, 60), rs.randint(0, 60), rs.uniform(0.01, 1))
        sks.append(sk)
    ref_keep = pp.refine_skeleton(copy.deepcopy(sks)); ora_keep = O.refine_skeleton(sks)
    assert len(ref_keep) == len(ora_keep)
    for sc in (1, 2, 4, 8):
        r = pp.skeleton_to_box(copy.deepcopy(ref_keep), sc); o = O.skeleton_to_box(ora_keep, sc)
        assert np.array_equal(np.asarray(r, np.float64), np.asarray(o, np.float64))


def test_forward_dec_and_seg_match_reference_module(reference):
    KGnet, _, _ = reference
    sd = O.make_state_dict(seed=0)
    model = KGnet.resnet50(pretrained=False).eval()
    missing = model.load_state_dict(sd, strict=True)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 64, 64) - 0.5
    with torch.no_grad():
        ref = model.forward_dec(x)
    ora = O.forward_dec(sd, x)
    for s in range(4):
        for a, b in zip(ref[s], ora[s]):
            assert torch.equal(a, b)
    for a, b in zip(ref[4], ora[4]):
        assert torch.equal(a, b)
    boxes = [np.array([[4., 6., 40., 50., 0.9], [10., 10., 20., 22., 0.5], [0., 0., 2., 2., 0.1], [30., 30., 31., 31., 0.05]])]
    with torch.no_grad():
        rseg = model.forward_seg(ref[4], boxes)
    oseg = O.forward_seg(sd, ora[4], boxes)
    assert len(rseg[0][0]) == len(oseg[0][0]) == 3   # the 1x1 box is dropped (KGnet.py:253)
    for a, b in zip(rseg[0][0], oseg[0][0]):
        assert torch.equal(a, b)
    for a, b in zip(rseg[1][0], oseg[1][0]):
        assert torch.equal(a, b)


def _reference_instance_heat():
    """The reference's InstanceHeat class (test.py) without running its constructor (which downloads ImageNet weights)."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("kg_reference_test_py", "/root/reference/test.py")
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    try:
        sys.argv = ["test.py"]
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod.InstanceHeat


def test_post_processing_restatement_matches_reference(reference):
    """oracle.post_processing against test.py:127-157 (same cv2 in the same process): bit-identical."""
    import types
    IH = _reference_instance_heat()
    obj = IH.__new__(IH)
    rs = np.random.RandomState(0)
    args = types.SimpleNamespace(input_h=96, input_w=128, seg_thresh=0.5)
    patches = [torch.from_numpy(rs.rand(37, 56).astype(np.float32)), torch.from_numpy(rs.rand(20, 30).astype(np.float32)),
               torch.from_numpy(rs.rand(17, 23).astype(np.float32))]
    dets = [torch.Tensor([3.2, 4.7, 40.4, 60.6, 0.9]), torch.Tensor([10., 20., 30., 50., 0.5]), torch.Tensor([8., 9., 41., 70., 0.3])]
    preds = [[patches[:2], patches[2:]], [dets[:2], dets[2:]]]
    for iw, ih in ((128, 96), (200, 150), (61, 47)):
        ref = IH.post_processing(obj, args, preds, iw, ih)
        ora = O.post_processing(preds, 96, 128, iw, ih, 0.5)
        assert ref[0].dtype == ora[0].dtype and np.array_equal(ref[0], ora[0]) and np.array_equal(ref[1], ora[1])
    assert O.post_processing(None, 96, 128, 10, 10, 0.5) is None


def test_preprocess_restatement_matches_reference_lines(reference):
    """oracle.preprocess_image = test.py:91-92."""
    import cv2
    rs = np.random.RandomState(1)
    image = rs.randint(0, 256, size=(70, 90, 3)).astype(np.uint8)
    img_input = cv2.resize(image, (64, 48))
    ref = torch.FloatTensor(np.transpose(img_input.copy(), (2, 0, 1))).unsqueeze(0) / 255 - 0.5
    assert torch.equal(ref, O.preprocess_image(image, 48, 64))


def _random_instances(rs, n, H, W):
    bb = []
    for _ in range(n):
        x1 = rs.randint(0, W - 14); y1 = rs.randint(0, H - 14)
        x2 = min(x1 + rs.randint(12, 31), W - 1); y2 = min(y1 + rs.randint(12, 31), H - 1)
        bb.append([(x1, y1), (x2, y1), (x1, y2), (x2, y2), ((x1 + x2) / 2, (y1 + y2) / 2)])
    return np.asarray(bb, np.float32).reshape(-1, 5, 2)


def test_ground_truth_encoder_restatement_matches_reference(reference):
    """oracle.encode_ground_truth against preprocessing.get_ground_truth + the concat of dataset_base.py:99-102 (bit-exact,
    including overlapping instance windows and keypoints at the image border)."""
    import preprocessing as P        # the reference's (conftest put /root/reference on sys.path)
    rs = np.random.RandomState(0)
    for trial in range(8):
        H, W = (48, 64) if trial < 5 else (40, 40)
        bb = _random_instances(rs, int(rs.randint(0, 8)), H, W)
        kp, sh, mid = P.get_ground_truth(bb, H, W, 5)
        ref = np.concatenate((kp, np.transpose(sh, (2, 0, 1)), np.transpose(mid, (2, 0, 1))), 0).astype(np.float32)
        assert np.array_equal(ref, O.encode_ground_truth(bb, H, W)), trial


def test_loss_restatements_match_reference_modules(reference):
    """oracle.detection_loss / seg_loss against loss.DetectionLossAll and seg_loss.SEG_loss (bit-exact on CPU)."""
    import warnings
    import loss as RL
    import seg_loss as RS
    torch.manual_seed(0)
    N, H, W = 2, 24, 32
    pr = [torch.rand(N, 5, H, W), torch.randn(N, 10, H, W), torch.randn(N, 40, H, W)]
    gt = torch.zeros(N, 55, H, W)
    gt[:, :5] = (torch.rand(N, 5, H, W) > 0.8).float(); gt[:, 5:] = torch.randn(N, 50, H, W)
    assert torch.equal(RL.DetectionLossAll(5)(pr, gt), O.detection_loss(pr, gt)[0])
    rs = np.random.RandomState(0)
    gt_masks = [rs.rand(3, H, W).round().astype(np.float32), rs.rand(2, H, W).round().astype(np.float32)]
    gt_boxes = [np.array([[2, 3, 14, 20, 1], [10, 10, 22, 30, 1], [0, 0, 5, 5, 1]], np.float32), np.array([[4, 4, 20, 28, 1], [1, 1, 3, 3, 1]], np.float32)]
    patches = [[torch.rand(12, 17) * 0.98 + 0.01, torch.rand(11, 19) * 0.98 + 0.01], [torch.rand(16, 24) * 0.98 + 0.01]]
    dets = [[torch.Tensor([2.2, 3.1, 14.4, 20.3, 0.9]), torch.Tensor([10.6, 10.2, 21.7, 29.9, 0.8])], [torch.Tensor([4, 4, 20, 28, 0.7])]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = RS.SEG_loss(H, W)([patches, dets], gt_masks, gt_boxes)
    assert torch.equal(ref, O.seg_loss([patches, dets], gt_masks, gt_boxes, H, W))
    assert O.seg_loss([[[]], [[]]], [gt_masks[0]], [gt_boxes[0]], H, W) is None


def test_loss_gradients_of_the_restatements_match_the_reference_modules(reference):
    """The device gradients (tests/test_train_side_gpu.py) are checked against torch autograd of the ORACLE losses: pin that
    autograd of the oracle equals autograd of the reference's own modules (`loss.backward()`, train.py:150)."""
    import warnings
    import loss as RL
    import seg_loss as RS
    torch.manual_seed(1)
    N, H, W = 2, 24, 32
    base = [torch.rand(N, 5, H, W) * 0.98 + 0.01, torch.randn(N, 10, H, W), torch.randn(N, 40, H, W)]
    gt = torch.zeros(N, 55, H, W)
    gt[:, :5] = (torch.rand(N, 5, H, W) > 0.8).float(); gt[:, 5:] = torch.randn(N, 50, H, W)
    a = [t.clone().requires_grad_(True) for t in base]
    b = [t.clone().requires_grad_(True) for t in base]
    RL.DetectionLossAll(5)(a, gt).backward()
    O.detection_loss(b, gt)[0].backward()
    for x, y in zip(a, b):
        assert torch.equal(x.grad, y.grad)
    rs = np.random.RandomState(0)
    gt_masks = [rs.rand(3, H, W).round().astype(np.float32), rs.rand(2, H, W).round().astype(np.float32)]
    gt_boxes = [np.array([[2, 3, 14, 20, 1], [10, 10, 22, 30, 1], [0, 0, 5, 5, 1]], np.float32), np.array([[4, 4, 20, 28, 1], [1, 1, 3, 3, 1]], np.float32)]
    shapes = [[(12, 17), (11, 19)], [(16, 24)]]
    dets = [[torch.Tensor([2.2, 3.1, 14.4, 20.3, 0.9]), torch.Tensor([10.6, 10.2, 21.7, 29.9, 0.8])], [torch.Tensor([4, 4, 20, 28, 0.7])]]
    pbase = [[torch.rand(s) * 0.98 + 0.01 for s in per] for per in shapes]
    pa = [[t.clone().requires_grad_(True) for t in per] for per in pbase]
    pb = [[t.clone().requires_grad_(True) for t in per] for per in pbase]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        RS.SEG_loss(H, W)([pa, dets], gt_masks, gt_boxes).backward()
    O.seg_loss([pb, dets], gt_masks, gt_boxes, H, W).backward()
    for per_a, per_b in zip(pa, pb):
        for x, y in zip(per_a, per_b):
            assert (x.grad is None) == (y.grad is None)
            if x.grad is not None:
                assert torch.equal(x.grad, y.grad)
