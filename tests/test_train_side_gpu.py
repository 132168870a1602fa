"""GPU parity of the callers either side of the inference path that SURVEY.md 8f ranks "next": the ground-truth encoder
(preprocessing.py:45-118) and the loss forward passes of the validation loop (loss.py, seg_loss.py; train.py:165-177)."""
import numpy as np
import pytest
import torch

from oracle import kg_oracle as O

pytestmark = pytest.mark.gpu


def _instances(rs, n, H, W, lo=12, hi=31):
    bb = []
    for _ in range(n):
        x1 = rs.randint(0, W - lo - 2); y1 = rs.randint(0, H - lo - 2)
        x2 = min(x1 + rs.randint(lo, hi), W - 1); y2 = min(y1 + rs.randint(lo, hi), H - 1)
        bb.append([(x1, y1), (x2, y1), (x1, y2), (x2, y2), ((x1 + x2) / 2, (y1 + y2) / 2)])
    return np.asarray(bb, np.float32).reshape(-1, 5, 2)


def test_ground_truth_encoder_bit_exact():
    """Batched device encoder against the oracle (itself pinned bit-exact against preprocessing.get_ground_truth): empty
    images, overlapping windows, keypoints on the border, 4 scales like dataset_base.py:90-97."""
    from kg_instance_segmentation_b200 import preprocessing
    rs = np.random.RandomState(1)
    for H, W in ((64, 96), (128, 128), (40, 40)):
        boxes = [_instances(rs, n, H, W) for n in (0, 1, 5, 17)]
        boxes[2][0, :, :] = [(0, 0), (W - 1, 0), (0, H - 1), (W - 1, H - 1), ((W - 1) / 2, (H - 1) / 2)]      # image-sized instance
        gt = preprocessing.encode_ground_truth_batch(boxes, H, W).cpu().numpy()
        for b, bb in enumerate(boxes):
            assert np.array_equal(gt[b], O.encode_ground_truth(bb, H, W)), (H, W, b)
    kp, sh, mid = preprocessing.get_ground_truth(boxes[1], 40, 40, 5)          # the reference's per-image return convention
    ref = O.encode_ground_truth(boxes[1], 40, 40)
    assert kp.shape == (5, 40, 40) and sh.shape == (40, 40, 10) and mid.shape == (40, 40, 40)
    assert np.array_equal(kp, ref[:5]) and np.array_equal(sh.transpose(2, 0, 1), ref[5:15]) and np.array_equal(mid.transpose(2, 0, 1), ref[15:])


def test_ground_truth_encoder_dense_1024():
    """cfg-4 sized input: 500 instances on a 1024x1024 map in one launch."""
    from kg_instance_segmentation_b200 import preprocessing
    rs = np.random.RandomState(2)
    bb = _instances(rs, 500, 1024, 1024, 16, 41)
    gt = preprocessing.encode_ground_truth_batch([bb], 1024, 1024)[0].cpu().numpy()
    assert np.array_equal(gt, O.encode_ground_truth(bb, 1024, 1024))


def test_detection_loss_matches_oracle():
    """DetectionLossAll.forward: fp32 torch reductions on the oracle side, fp64 accumulation here: rtol 2e-6."""
    from kg_instance_segmentation_b200 import loss
    torch.manual_seed(0)
    for N, H, W in ((2, 64, 96), (1, 128, 128), (3, 16, 16)):
        pr = [torch.rand(N, 5, H, W) * 0.98 + 0.01, torch.randn(N, 10, H, W) * 3, torch.randn(N, 40, H, W) * 20]
        gt = torch.zeros(N, 55, H, W)
        gt[:, :5] = (torch.rand(N, 5, H, W) > 0.9).float(); gt[:, 5:] = torch.randn(N, 50, H, W) * 5
        crit = loss.DetectionLossAll(5)
        got = crit([t.cuda() for t in pr], gt.cuda())
        ref, kp, sh, mid = O.detection_loss(pr, gt)
        assert got.shape == () and got.is_cuda
        np.testing.assert_allclose(float(got), float(ref), rtol=2e-6)
        np.testing.assert_allclose(crit.last_terms.cpu().numpy(), [float(kp), float(sh), float(mid)], rtol=2e-6)
    # saturated predictions: the -100 clamp of F.binary_cross_entropy, and an all-zero target (denominator 1e-10)
    pr = [torch.zeros(1, 5, 8, 8), torch.ones(1, 10, 8, 8), torch.ones(1, 40, 8, 8)]
    gt = torch.zeros(1, 55, 8, 8); gt[0, 0, 0, 0] = 1.0
    got = loss.DetectionLossAll(5)([t.cuda() for t in pr], gt.cuda())
    np.testing.assert_allclose(float(got), float(O.detection_loss(pr, gt)[0]), rtol=2e-6)
    gt.zero_()
    got = loss.DetectionLossAll(5)([t.cuda() for t in pr], gt.cuda())
    np.testing.assert_allclose(float(got), float(O.detection_loss(pr, gt)[0]), rtol=2e-6, atol=1e-12)


def test_seg_loss_matches_oracle_on_forward_seg_output():
    """SEG_loss.forward on the real forward_seg output (patches are windows of the packed atlas) and on loose tensors."""
    from kg_instance_segmentation_b200 import KGnet, seg_loss
    sd = O.make_state_dict(seed=0)
    sd["seg_head.2.weight"] = sd["seg_head.2.weight"] * 0.05
    m = KGnet.resnet50(pretrained=False, precision="exact")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    torch.manual_seed(3)
    H = W = 96
    x = torch.rand(2, 3, H, W) - 0.5
    gt_boxes = [np.array([[4., 6., 50., 60., 1.], [40., 30., 90., 88., 1.]], np.float32), np.array([[10., 12., 70., 64., 1.]], np.float32)]
    rs = np.random.RandomState(0)
    gt_masks = [rs.rand(2, H, W).round().astype(np.float32), rs.rand(1, H, W).round().astype(np.float32)]
    det_boxes = [np.array([[5., 7., 49., 61., 0.9], [42., 31., 88., 86., 0.8], [0., 0., 10., 10., 0.3]]), np.array([[11., 12., 69., 66., 0.7]])]
    out = m.forward_dec(x.cuda())
    preds = m.forward_seg(out[4], det_boxes)
    got = seg_loss.SEG_loss(H, W)(preds, gt_masks, gt_boxes)
    ref = O.seg_loss(preds, gt_masks, gt_boxes, H, W)
    assert got is not None and ref is not None
    np.testing.assert_allclose(float(got), float(ref), rtol=5e-6)
    loose = [[[p.clone() for p in per] for per in preds[0]], preds[1]]        # plain list: patches are gathered into one buffer
    np.testing.assert_allclose(float(seg_loss.SEG_loss(H, W)(loose, gt_masks, gt_boxes)), float(ref), rtol=5e-6)
    far = [np.array([[60., 60., 90., 90., 1.]], np.float32), np.array([[80., 80., 95., 95., 1.]], np.float32)]
    assert seg_loss.SEG_loss(H, W)(preds, [gt_masks[0][:1], gt_masks[1]], far) is None and O.seg_loss(preds, [gt_masks[0][:1], gt_masks[1]], far, H, W) is None


def test_detection_loss_gradient_matches_autograd_of_the_oracle():
    """`loss.backward()` (train.py:150): gradients with respect to the three prediction tensors against torch autograd of the
    oracle's restatement on the CPU.  fp32 elementwise formulas; the denominators are fp64 sums here: rtol 1e-5."""
    from kg_instance_segmentation_b200 import loss
    torch.manual_seed(1)
    for N, H, W in ((2, 32, 48), (1, 64, 64)):
        pr = [torch.rand(N, 5, H, W) * 0.98 + 0.01, torch.randn(N, 10, H, W) * 3, torch.randn(N, 40, H, W) * 20]
        gt = torch.zeros(N, 55, H, W)
        gt[:, :5] = (torch.rand(N, 5, H, W) > 0.9).float(); gt[:, 5:] = torch.randn(N, 50, H, W) * 5
        gt[:, 5:7][:, :, :4] = pr[1][:, :2, :4]                     # exact ties: sign(0) = 0 like torch.abs' backward
        ref_in = [t.clone().requires_grad_(True) for t in pr]
        (O.detection_loss(ref_in, gt)[0] * 3.0).backward()
        dev_in = [t.clone().cuda().requires_grad_(True) for t in pr]
        got = loss.DetectionLossAll(5)(dev_in, gt.cuda())
        assert got.requires_grad
        (got * 3.0).backward()
        for a, b in zip(dev_in, ref_in):
            scale = float(b.grad.abs().max())
            np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-6 * scale)
    # saturated keypoint predictions: PyTorch clamps the BCE-backward denominator at 1e-12
    pr = [torch.tensor([0.0, 1.0, 1e-8, 0.5, 1 - 1e-7]).reshape(1, 5, 1, 1).repeat(1, 1, 2, 2), torch.ones(1, 10, 2, 2), torch.ones(1, 40, 2, 2)]
    gt = torch.zeros(1, 55, 2, 2); gt[0, 1] = 1.0; gt[0, 2, 0, 0] = 1.0
    ref_in = [t.clone().requires_grad_(True) for t in pr]
    O.detection_loss(ref_in, gt)[0].backward()
    dev_in = [t.clone().cuda().requires_grad_(True) for t in pr]
    loss.DetectionLossAll(5)(dev_in, gt.cuda()).backward()
    for a, b in zip(dev_in, ref_in):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-30)
    # predictions that do not require grad: a plain tensor comes back (validation loop)
    assert not loss.DetectionLossAll(5)([t.cuda() for t in pr], gt.cuda()).requires_grad


def test_seg_loss_gradient_matches_autograd_of_the_oracle():
    """SEG_loss on reference-style lists of patch tensors that require grad: d loss / d patch against torch autograd of the oracle
    (a patch matched with two ground-truth objects accumulates both terms)."""
    from kg_instance_segmentation_b200 import seg_loss
    H = W = 80
    rs = np.random.RandomState(5)
    torch.manual_seed(5)
    gt_boxes = [np.array([[4., 6., 50., 60., 1.], [6., 8., 52., 58., 1.]], np.float32), np.array([[10., 12., 70., 64., 1.]], np.float32)]
    gt_masks = [rs.rand(2, H, W).round().astype(np.float32), rs.rand(1, H, W).round().astype(np.float32)]
    dets = [[np.array([5., 7., 49., 61., 0.9], np.float32), np.array([0., 0., 10., 10., 0.3], np.float32)], [np.array([11., 12., 69., 66., 0.7], np.float32)]]
    shapes = [[(44, 54), (10, 10)], [(58, 54)]]
    base = [[torch.rand(s) * 0.96 + 0.02 for s in per] for per in shapes]
    ref_p = [[t.clone().requires_grad_(True) for t in per] for per in base]
    ref = O.seg_loss([ref_p, dets], gt_masks, gt_boxes, H, W)
    ref.backward()
    dev_p = [[t.clone().cuda().requires_grad_(True) for t in per] for per in base]
    got = seg_loss.SEG_loss(H, W)([dev_p, dets], gt_masks, gt_boxes)
    np.testing.assert_allclose(float(got), float(ref), rtol=5e-6)
    got.backward()
    for per_d, per_r in zip(dev_p, ref_p):
        for a, b in zip(per_d, per_r):
            if b.grad is None:                      # the unmatched patch
                assert a.grad is None or float(a.grad.abs().max()) == 0.0
            else:
                np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=2e-5, atol=1e-9)


def test_validation_step_flow():
    """train.py:165-177 `validating` with the three imports swapped: forward(x, boxes) -> 4 x DetectionLossAll + SEG_loss."""
    from kg_instance_segmentation_b200 import KGnet, loss, preprocessing, seg_loss
    sd = O.make_state_dict(seed=0)
    m = KGnet.resnet50(pretrained=False, precision="exact")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    H = W = 64
    torch.manual_seed(5)
    x = torch.rand(1, 3, H, W) - 0.5
    rs = np.random.RandomState(4)
    inst = _instances(rs, 2, H, W, 14, 30)
    gts = [preprocessing.encode_ground_truth_batch([np.floor(inst / s)], H // s, W // s) for s in (1, 2, 4, 8)]
    bboxes_c0 = [np.array([[b[0, 1], b[0, 0], b[3, 1], b[3, 0], 1.] for b in inst], np.float32)]
    masks = [np.stack([np.pad(np.ones((int(b[3, 1] - b[0, 1]), int(b[3, 0] - b[0, 0])), np.float32),
                              ((int(b[0, 1]), H - int(b[3, 1])), (int(b[0, 0]), W - int(b[3, 0])))) for b in inst])]
    pr_c0, pr_c1, pr_c2, pr_c3, predictions = m(x.cuda(), bboxes_c0)
    loss_dec, loss_seg = loss.DetectionLossAll(5), seg_loss.SEG_loss(H, W)
    loss1 = loss_dec(pr_c0, gts[0]) + loss_dec(pr_c1, gts[1]) + loss_dec(pr_c2, gts[2]) + loss_dec(pr_c3, gts[3])
    loss2 = loss_seg(predictions, masks, bboxes_c0)
    total = float((loss1 + loss2).item())
    ref_out = O.forward_dec(sd, x)
    ref1 = sum(O.detection_loss(ref_out[s], torch.from_numpy(O.encode_ground_truth(np.floor(inst / sc), H // sc, W // sc))[None])[0]
               for s, sc in enumerate((1, 2, 4, 8)))
    ref2 = O.seg_loss(O.forward_seg(sd, ref_out[4], bboxes_c0), masks, bboxes_c0, H, W)
    np.testing.assert_allclose(total, float(ref1 + ref2), rtol=2e-3)


def test_training_step_through_an_autograd_network():
    """train.py:145-154 with the loss modules swapped: an autograd network (here a small torch conv net standing in for the
    reference's nn.Module) -> DetectionLossAll -> loss.backward() -> optimizer step.  The parameter gradients through this library's
    fused loss kernels must equal those through the oracle's torch loss, and an SGD step must lower the loss."""
    from kg_instance_segmentation_b200 import loss
    torch.manual_seed(0)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Conv2d(3, 16, 3, padding=1)
            self.kp = torch.nn.Conv2d(16, 5, 3, padding=1)
            self.sh = torch.nn.Conv2d(16, 10, 3, padding=1)
            self.mid = torch.nn.Conv2d(16, 40, 3, padding=1)

        def forward(self, x):
            f = torch.relu(self.body(x))
            return [torch.sigmoid(self.kp(f)), self.sh(f), self.mid(f)]

    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        net_a, net_b = Tiny().cuda(), Tiny().cuda()
        net_b.load_state_dict(net_a.state_dict())
        x = torch.rand(2, 3, 48, 64, device="cuda") - 0.5
        gt = torch.zeros(2, 55, 48, 64, device="cuda")
        gt[:, :5] = (torch.rand(2, 5, 48, 64, device="cuda") > 0.9).float(); gt[:, 5:] = torch.randn(2, 50, 48, 64, device="cuda") * 3
        crit = loss.DetectionLossAll(5)
        la = crit(net_a(x), gt)
        la.backward()
        lb = O.detection_loss(net_b(x), gt)[0]
        lb.backward()
        np.testing.assert_allclose(float(la), float(lb), rtol=2e-6)
        for (name, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
            scale = float(pb.grad.abs().max())
            np.testing.assert_allclose(pa.grad.cpu().numpy(), pb.grad.cpu().numpy(), rtol=2e-4, atol=2e-6 * scale, err_msg=name)
        opt = torch.optim.SGD(net_a.parameters(), lr=0.05)
        first = float(la)
        for _ in range(5):
            opt.zero_grad()
            l = crit(net_a(x), gt)
            l.backward()
            opt.step()
        assert float(crit(net_a(x), gt)) < first
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_fused_adam_matches_torch_adam():
    """optim.Adam (ONE launch for all tensors) against torch.optim.Adam on the CPU (train.py:71,154): parameters and both moment
    buffers after several steps, an lr scheduler in between, one parameter that skips a step (its own step count must not advance),
    tensors larger than one 65 536-element chunk."""
    from kg_instance_segmentation_b200 import optim
    torch.manual_seed(0)
    shapes = [(64, 3, 7, 7), (64,), (300, 257), (1,), (70000,)]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    dev_p = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref_p]
    ref_o = torch.optim.Adam(ref_p, lr=1e-2)
    dev_o = optim.Adam(dev_p, lr=1e-2)
    ref_s = torch.optim.lr_scheduler.ExponentialLR(ref_o, gamma=0.96)
    dev_s = torch.optim.lr_scheduler.ExponentialLR(dev_o, gamma=0.96)
    for it in range(6):
        for i, (a, b) in enumerate(zip(ref_p, dev_p)):
            if i == 3 and it == 2:
                a.grad = None; b.grad = None                       # this parameter sits out one step
                continue
            g = torch.randn(shapes[i]) * (10.0 ** (it - 3))        # gradients over six orders of magnitude
            a.grad = g.clone(); b.grad = g.clone().cuda()
        ref_o.step(); dev_o.step()
        ref_s.step(); dev_s.step()
        for i, (a, b) in enumerate(zip(ref_p, dev_p)):
            # agreement to an ulp or two of the tensor's magnitude (measured: <= 1.2e-7 relative to the largest element; ATen's vectorised
            # CPU kernels contract some multiply-adds, the device kernel follows the scalar operation order)
            for got, want, what in ((b.detach(), a.detach(), "param"), (dev_o.state[b]["exp_avg"], ref_o.state[a]["exp_avg"], "exp_avg"),
                                    (dev_o.state[b]["exp_avg_sq"], ref_o.state[a]["exp_avg_sq"], "exp_avg_sq")):
                scale = float(want.abs().max())
                np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-6, atol=5e-7 * scale + 1e-30,
                                           err_msg=f"{what} {i} step {it}")
    assert dev_o.state[dev_p[3]]["step"] == 5 and dev_o.state[dev_p[0]]["step"] == 6
    cpu_p = torch.nn.Parameter(torch.zeros(3))
    cpu_p.grad = torch.zeros(3)
    with pytest.raises(RuntimeError):
        optim.Adam([cpu_p], lr=1e-3).step()                    # no CPU fallback
