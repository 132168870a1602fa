"""GPU parity tests of the network path: single convs by shape class, forward_dec and forward_seg against the
oracle (torch fp32 functional ops on CPU) and the golden vectors generated from the reference.

Tolerances: the CUDA-core path (mode 0) and the 3-pass split-fp16 tensor-core path accumulate in fp32 in a
different order than oneDNN: rtol 2e-4 of the output scale.  The single-pass fp16 path rounds operands to 11
bits: 3e-3 of the output scale per layer.  End to end the contract is 1e-3 on the keypoint heatmaps."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _conv(x, w, b, stride, pad, relu, res, mode):
    from kg_instance_segmentation_b200 import _cabi
    L = _cabi.lib()
    N, Cin, H, W = x.shape
    Cout, _, R, S = w.shape
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    xd = x.cuda().contiguous()
    y = torch.empty(N, Cout, Ho, Wo, device="cuda")
    rd = res.cuda().contiguous() if res is not None else None
    wc = w.contiguous(); bc = b.contiguous() if b is not None else None
    _cabi.check(L.kg_conv2d_nchw(xd.data_ptr(), N, Cin, H, W, wc.data_ptr(), bc.data_ptr() if bc is not None else None, Cout, R, S,
                                 stride, pad, int(relu), rd.data_ptr() if rd is not None else None, mode, y.data_ptr(), None))
    return y.cpu()


def _case(seed, N, Cin, H, W, Cout, k, stride, relu, res):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    pad = k // 2
    y = F.conv2d(x, w, b, stride=stride, padding=pad)
    r = torch.randn_like(y) if res else None
    if r is not None:
        y = y + r
    if relu:
        y = F.relu(y)
    return x, w, b, pad, r, y


FFMA_CASES = [  # (N, Cin, H, W, Cout, k, stride, relu, res): the shape classes of SURVEY.md §8a
    (2, 3, 40, 36, 64, 3, 1, True, False),      # c0_conv.0 stem
    (2, 3, 40, 36, 64, 7, 2, True, False),      # conv1 7x7/s2
    (1, 64, 24, 20, 64, 3, 1, True, False),     # 3x3 64->64
    (2, 128, 16, 16, 128, 3, 2, True, False),   # layerN.0.conv2 stride 2
    (2, 256, 16, 12, 512, 1, 2, False, False),  # downsample 1x1/s2
    (1, 256, 8, 8, 64, 1, 1, True, True),       # 1x1 + residual + ReLU
    (1, 64, 12, 12, 5, 7, 1, False, False),     # head layer 2 (ragged Cout)
    (1, 64, 9, 7, 1, 3, 1, False, False),       # seg_head.2, odd sizes
]


@pytest.mark.parametrize("case", FFMA_CASES)
def test_conv_cuda_core_path(case):
    x, w, b, pad, r, y = _case(1, *case)
    got = _conv(x, w, b, case[6], pad, case[7], r, 0)
    scale = float(y.abs().max())
    assert float((got - y).abs().max()) <= 2e-4 * scale + 1e-5


TC_CASES = [
    (2, 64, 32, 32, 64, 3, 1, True, False),     # c1_up_conv class (N = 64)
    (1, 64, 16, 48, 64, 3, 1, True, False),     # W not a power of two: partial tiles, OOB rows
    (2, 64, 16, 16, 256, 1, 1, False, True),    # bottleneck conv3 + residual
    (1, 256, 16, 16, 64, 1, 1, True, False),    # bottleneck conv1
    (1, 128, 16, 16, 128, 3, 1, True, False),
    (1, 1024, 8, 8, 512, 3, 1, True, False),    # c4_up_conv class (N tile 256 x 2, K = 9216)
    (1, 64, 24, 24, 192, 7, 1, True, False),    # fused first-layer heads c0/c1 (N = 192)
    (1, 256, 8, 8, 768, 7, 1, True, False),     # fused first-layer heads c2 (3 N tiles)
    (2, 64, 4, 4, 64, 3, 1, False, False),      # tiny map: tile taller than the image
    (3, 64, 8, 8, 64, 1, 1, False, False),      # odd number of M tiles
    (1, 64, 6, 128, 64, 3, 1, True, False),     # W >= 128: strip mode (taps of a filter row share one activation strip)
    (2, 64, 5, 256, 192, 7, 1, True, False),    # strip mode, 7x7, two tiles per row
    (1, 128, 3, 200, 128, 3, 1, False, False),  # strip mode with a partial last tile
    (1, 64, 4, 128, 16, 7, 1, False, False),    # strip mode, narrow N (second-layer heads class)
]


TC_STRIDE2_CASES = [
    (2, 128, 16, 16, 128, 3, 2, True, False),   # layer2.0.conv2
    (1, 256, 32, 32, 512, 1, 2, False, False),  # layer2.0.downsample
    (1, 256, 16, 24, 256, 3, 2, True, False),
]


@pytest.mark.parametrize("case", TC_STRIDE2_CASES)
def test_conv_tensor_core_stride2(case):
    x, w, b, pad, r, y = _case(4, *case)
    got = _conv(x, w, b, 2, pad, case[7], r, 3)
    scale = float(y.abs().max())
    err = float((got - y).abs().max())
    assert err <= 2e-4 * scale + 1e-5, f"max err {err}"


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("mode", [3, 2, 1])
def test_conv_tensor_core_path(case, mode):
    from kg_instance_segmentation_b200 import _cabi
    if not _cabi.lib().kg_tc_available():
        pytest.fail("tcgen05 path unavailable: " + _cabi.lib().kg_tc_status().decode())
    x, w, b, pad, r, y = _case(2, *case)
    got = _conv(x, w, b, 1, pad, case[7], r, mode)
    scale = float(y.abs().max())
    tol = (2e-4 if mode == 3 else 3e-3) * scale + 1e-5      # 2-pass: split activations x fp16 weights (11-bit weight rounding)
    err = float((got - y).abs().max())
    assert err <= tol, f"max err {err} > {tol}"


HEADS_L2_CASES = [  # (N, Cin, H, W): second-layer head convs through the row-GEMM + shift-add kernel (tc_shift.cu)
    (1, 64, 9, 512),     # row mode, 4 tiles per row: carries across tile edges
    (2, 64, 5, 256),     # row mode, 2 tiles per row, 2 images
    (1, 256, 6, 128),    # row mode, one tile per row, 4 K chunks
    (2, 128, 10, 64),    # two image rows per tile
    (1, 64, 7, 32),      # four rows per tile, H not a multiple of the tile height
    (1, 64, 16, 16),
    (3, 64, 8, 8),       # tile taller than the image
    (1, 512, 4, 64),     # c3 class: 8 K chunks
]


@pytest.mark.parametrize("case", HEADS_L2_CASES)
def test_heads_l2_shift_kernel(case):
    from kg_instance_segmentation_b200 import _cabi
    L = _cabi.lib()
    N, Cin, H, W = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 3 * Cin, H, W, generator=g)
    ws = [torch.randn(co, Cin, 7, 7, generator=g) * (2.0 / (Cin * 49)) ** 0.5 for co in (5, 10, 40)]
    bs = [torch.randn(co, generator=g) * 0.1 for co in (5, 10, 40)]
    ref = [F.conv2d(x[:, h * Cin:(h + 1) * Cin], ws[h], bs[h], padding=3) for h in range(3)]
    ref[0] = torch.sigmoid(ref[0])
    xd = x.cuda().contiguous()
    ys = [torch.full((N, co, H, W), float("nan"), device="cuda") for co in (5, 10, 40)]
    wp = (C.c_void_p * 3)(*[w.data_ptr() for w in ws]); bp = (C.c_void_p * 3)(*[b.data_ptr() for b in bs])
    yp = (C.c_void_p * 3)(*[y.data_ptr() for y in ys])
    _cabi.check(L.kg_heads_l2_nchw(xd.data_ptr(), N, Cin, H, W, wp, bp, yp, None))
    for h in range(3):
        got = ys[h].cpu()
        assert not torch.isnan(got).any(), f"head {h}: unwritten outputs"
        scale = float(ref[h].abs().max())
        err = float((got - ref[h]).abs().max())
        assert err <= 3e-3 * scale + 1e-5, f"head {h}: max err {err} (scale {scale})"


SHIFT_CONV_CASES = [  # (N, Cin, H, W, Cout): 3x3 convs through the row-GEMM + shift-add kernel (tc_shift.cu)
    (1, 64, 7, 512, 64),     # c1_up_conv / c0_conv.2 class, 4 tiles per row
    (2, 256, 5, 256, 64),    # c2_up_conv class: 4 K chunks, 2 tiles per row
    (2, 64, 9, 128, 64),     # layer1 conv2 class
    (1, 64, 12, 64, 64),     # two rows per tile
    (3, 64, 8, 8, 64),
    (1, 64, 6, 256, 1),      # seg_head.2: one output channel, fp32 output
    (1, 64, 10, 32, 1),
]


@pytest.mark.parametrize("case", SHIFT_CONV_CASES)
@pytest.mark.parametrize("mode", [13, 12, 11])
def test_conv_shift_kernel(case, mode):
    N, Cin, H, W, Cout = case
    x, w, b, pad, r, y = _case(6, N, Cin, H, W, Cout, 3, 1, Cout != 1, False)
    if mode == 12 and Cout == 1:
        pytest.skip("the one-channel kernel has no 2-pass variant")
    got = _conv(x, w, b, 1, pad, Cout != 1, None, mode)
    scale = float(y.abs().max())
    tol = (2e-4 if mode == 13 else 3e-3) * scale + 1e-5
    err = float((got - y).abs().max())
    assert err <= tol, f"max err {err} > {tol}"


TWO_CTA_CASES = [  # first-layer head class through the CTA-pair (cta_group::2) kernel: single pass, hi-plane output
    (2, 64, 6, 256, 192, 7, 1, True, False),    # c0 / c1 class: strip mode, N = 192, 2 tiles per row
    (1, 64, 5, 128, 192, 7, 1, True, False),    # odd number of pixel tiles: the peer CTA of the last pair idles
    (1, 256, 4, 128, 768, 7, 1, True, False),   # c2 class: N tile 256 x 3, 4 K chunks
    (2, 128, 8, 64, 256, 3, 1, True, False),    # two image rows per tile (no strip)
    (1, 64, 16, 16, 192, 7, 1, True, False),    # small map
]


@pytest.mark.parametrize("case", TWO_CTA_CASES)
def test_conv_cta_pair_kernel(case):
    x, w, b, pad, r, y = _case(8, *case)
    got = _conv(x, w, b, 1, pad, case[7], r, 21)
    scale = float(y.abs().max())
    err = float((got - y).abs().max())
    assert err <= 3e-3 * scale + 1e-5, f"max err {err}"


def _model(precision):
    from kg_instance_segmentation_b200 import KGnet
    sd = O.make_state_dict(seed=0)
    m = KGnet.resnet50(pretrained=False, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("precision,kp_tol,off_tol", [("reference", 2e-5, 2e-3), ("exact", 1e-4, 2e-3), ("fast", 1e-3, 5e-2)])
def test_forward_dec_matches_golden(precision, kp_tol, off_tol):
    g = np.load(os.path.join(G, "forward_64_seed0.npz"))
    m, _ = _model(precision)
    out = m.forward_dec(torch.from_numpy(g["x"]).cuda())
    for s in range(4):
        kp, sh, mid = (t.cpu().numpy() for t in out[s])
        assert kp.shape == g[f"ref_kp{s}"].shape
        assert np.abs(kp - g[f"ref_kp{s}"]).max() <= kp_tol, (s, np.abs(kp - g[f"ref_kp{s}"]).max())
        for got, name in ((sh, "short"), (mid, "mid")):
            ref = g[f"ref_{name}{s}"]
            assert np.abs(got - ref).max() <= off_tol * max(1.0, np.abs(ref).max()), (s, name, np.abs(got - ref).max())
    for l in range(5):
        ref = g[f"ref_c{l}"].astype(np.float32)
        got = out[4][l].cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("precision", ["reference", "fast"])
def test_forward_seg_matches_golden(precision):
    g = np.load(os.path.join(G, "forward_64_seed0.npz"))
    m, _ = _model(precision)
    out = m.forward_dec(torch.from_numpy(g["x"]).cuda())
    boxes = [g["boxes0"], g["boxes1"]]
    seg = m.forward_seg(out[4], boxes)
    for i in range(2):
        n_ref = sum(1 for k in g.files if k.startswith(f"ref_mask{i}_"))
        assert len(seg[0][i]) == len(seg[1][i]) == n_ref
        for j in range(n_ref):
            ref = g[f"ref_mask{i}_{j}"]
            got = seg[0][i][j].cpu().numpy()
            assert got.shape == ref.shape
            # masks: CUDA-core path 2e-3.  "fast" runs atlas level 0 of the mask branch in single-pass fp16; the golden
            # weights are raw Kaiming init, whose mask logits reach |z| ~ 50, so a 1e-3 relative logit error shows up as
            # ~1e-2 on the sigmoid: 2e-2 here, and the thresholded mask (test.py:149, seg_thresh 0.5) must agree wherever
            # the reference is not within that margin of the threshold.  The calibrated-logit case is tested below at 3e-3.
            tol = 2e-3 if precision == "reference" else 2e-2
            assert np.abs(got - ref).max() <= tol, (i, j, np.abs(got - ref).max())
            sure = np.abs(ref - 0.5) > tol
            assert np.array_equal((got >= 0.5)[sure], (ref >= 0.5)[sure])
            assert np.array_equal(seg[1][i][j].numpy(), g[f"ref_det{i}_{j}"])


def test_forward_seg_fast_calibrated_logits():
    """Mask branch in precision "fast" against the oracle with O(1) mask logits (seg_head.2 scaled like the calibrated
    keypoint heads): 3e-3 on the sigmoid output."""
    from kg_instance_segmentation_b200 import KGnet
    sd = O.make_state_dict(seed=0)
    sd["seg_head.2.weight"] = sd["seg_head.2.weight"] * 0.05
    m = KGnet.resnet50(pretrained=False, precision="fast")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    torch.manual_seed(7)
    x = torch.rand(2, 3, 64, 64) - 0.5
    ref = O.forward_dec(sd, x)
    boxes = [np.array([[3., 2., 60., 58., 0.9], [20., 10., 40., 44., 0.6]]), np.array([[8., 8., 30., 30., 0.7]])]
    rseg = O.forward_seg(sd, ref[4], boxes)
    out = m.forward_dec(x.cuda())
    seg = m.forward_seg(out[4], boxes)
    worst = 0.0
    for i in range(2):
        assert len(seg[0][i]) == len(rseg[0][i])
        for a, b in zip(seg[0][i], rseg[0][i]):
            assert a.shape == b.shape
            worst = max(worst, float((a.cpu() - b).abs().max()))
    assert worst <= 3e-3, worst


def test_forward_seg_on_external_features_and_empty_boxes():
    m, sd = _model("reference")
    torch.manual_seed(3)
    x = torch.rand(2, 3, 96, 80) - 0.5
    ref = O.forward_dec(sd, x)
    boxes = [np.array([[2., 3., 70., 60., 0.9], [30., 30., 36., 37., 0.4]]), []]
    rseg = O.forward_seg(sd, ref[4], boxes)
    seg = m.forward_seg([t.cuda() for t in ref[4]], boxes)       # features that did NOT come from our forward_dec
    assert len(seg[0][0]) == len(rseg[0][0]) and seg[0][1] == [] and seg[1][1] == []
    for a, b in zip(seg[0][0], rseg[0][0]):
        assert a.shape == b.shape and float((a.cpu() - b).abs().max()) <= 1e-4
    out = m(x.cuda(), boxes)                                      # ResNet.forward (KGnet.py:269-272)
    assert len(out) == 5 and len(out[4][0][0]) == len(rseg[0][0])
    none = m.forward_seg(out[3] if False else m.forward_dec(x.cuda())[4], [[], []])
    assert none == [[[], []], [[], []]]


def test_forward_dec_batch_consistency_and_sizes():
    """Non-square input and batch > 1: every image of a batch equals the same image run alone (fast precision)."""
    m, sd = _model("fast")
    torch.manual_seed(5)
    x = torch.rand(3, 3, 64, 96) - 0.5
    out = m.forward_dec(x.cuda())
    single = m.forward_dec(x[1:2].cuda())
    for s in range(4):
        for a, b in zip(out[s], single[s]):
            assert torch.equal(a[1:2], b)
    ref = O.forward_dec(sd, x)
    for s in range(4):
        assert float((out[s][0].cpu() - ref[s][0]).abs().max()) <= 1e-3


@pytest.mark.parametrize("precision,kp_tol,feat_tol", [("exact", 1e-4, 2e-5), ("fast", 1e-3, 2e-5)])
def test_forward_dec_u8_matches_oracle_and_float_path(precision, kp_tol, feat_tol):
    """uint8 NHWC input with the normalisation folded into the stem convs (kg_net_forward_dec_u8) against the oracle fed with the
    reference's `x / 255 - 0.5` (test.py:92) and against this library's own fp32-input path.  The c0 / c1 feature maps come straight
    out of the two stem kernels' consumers: their borders check the out-of-image tap value (-0.5 in the shifted integer domain)."""
    from kg_instance_segmentation_b200 import inference
    m, sd = _model(precision)
    rs = np.random.RandomState(7)
    img = rs.randint(0, 256, (2, 96, 80, 3)).astype(np.uint8)
    img[0, :3] = 255; img[0, -2:] = 0; img[1, :, :2] = 255; img[1, :, -3:] = 0       # extreme values on every border
    x_ref = torch.from_numpy(np.transpose(img, (0, 3, 1, 2)).copy()).float() / 255 - 0.5
    ref = O.forward_dec(sd, x_ref)
    d_img = torch.from_numpy(img).cuda()
    out = m.forward_dec_u8(d_img)
    flt = m.forward_dec(inference.preprocess_u8(d_img))
    for s in range(4):
        assert float((out[s][0].cpu() - ref[s][0]).abs().max()) <= kp_tol, s
        # the two input paths differ by the stems' rounding only (1e-6 on the heat maps); the single-pass fp16 heads of `fast`
        # amplify that to the size of their own rounding steps
        assert float((out[s][0] - flt[s][0]).abs().max()) <= (1e-4 if precision == "exact" else 5e-4), s
    for l in (0, 1):       # c0 (3x3 stem -> c0_conv.2 -> ...) and c1 features: relative to the map's magnitude
        scale = max(1.0, float(ref[4][l].abs().max()))
        assert float((out[4][l].cpu() - ref[4][l]).abs().max()) <= (feat_tol if precision == "exact" else 5e-3) * scale, l
    with pytest.raises(ValueError):
        m.forward_dec_u8(d_img.float())
    m.precision = "reference"
    with pytest.raises(RuntimeError):
        m.forward_dec_u8(d_img)


def test_workspace_is_placed_by_liveness():
    """The activation workspace reuses the planes of dead tensors: a bs32 / 512 x 512 `fast` pass needs well under half of the
    21 GB a consecutive layout takes (the parity tests of this file and of test_parity_full_gpu.py run on the reused layout)."""
    from kg_instance_segmentation_b200 import _cabi
    m, _ = _model("fast")
    m._sync_weights()
    nbytes = _cabi.lib().kg_net_workspace_bytes(m._handle, 32, 512, 512, 1)
    assert 4e9 < nbytes < 11e9, nbytes
