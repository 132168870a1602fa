"""Drop-in for the reference's KGnet.py (`resnet50()`, `ResNet.forward_dec / forward_seg / forward`).

The module is a torch `nn.Module` only as a PARAMETER CONTAINER: it registers the reference's 346 state-dict
tensors under the reference's names so that `load_state_dict(torch.load('end_model.pth'))`, `.to(device)` and
`.eval()` work unchanged (test.py:53,61,196-197).  All arithmetic of `forward_dec` / `forward_seg` runs in the
sm_100a kernels of libkgnet_b200.so (csrc/net.cu, net_kernels.cu, tc_conv.cu) through the C-ABI; there is no
torch-op or CPU fallback, and training-mode BatchNorm is not part of this path (eval-mode statistics are folded
into the convolutions when the weights are uploaded).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi

PRECISIONS = {"reference": 0, "ffma": 0, "fast": 1, "exact": 2}
_FEAT_C = (64, 64, 256, 512, 1024)
_HEADS = (("kp_head", 5), ("short_offset_head", 10), ("mid_offset_head", 40))


class _Conv(nn.Module):
    """weight [Cout,Cin,k,k] (+ bias): Kaiming-normal fan_out like KGnet.py:212-214, default Conv2d bias init."""

    def __init__(self, cin, cout, k, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")
        if bias:
            bound = 1.0 / math.sqrt(cin * k * k)
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class _BN(nn.Module):
    """BatchNorm2d state (KGnet.py:215-217: weight 1, bias 0); used in eval mode only."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


def _seq(**children):
    """Container whose children carry the numeric names of the reference's nn.Sequential slots."""
    m = nn.Module()
    for k, v in children.items():
        m.add_module(k.lstrip("_"), v)
    return m


class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, downsample):
        super().__init__()
        self.conv1, self.bn1 = _Conv(inplanes, planes, 1, False), _BN(planes)
        self.conv2, self.bn2 = _Conv(planes, planes, 3, False), _BN(planes)
        self.conv3, self.bn3 = _Conv(planes, planes * 4, 1, False), _BN(planes * 4)
        if downsample:
            self.downsample = _seq(_0=_Conv(inplanes, planes * 4, 1, False), _1=_BN(planes * 4))


class _Combination(nn.Module):
    def __init__(self, in_size, out_size, cat_size):
        super().__init__()
        self.up = _seq(_0=_Conv(in_size, out_size, 3, True))
        self.cat_conv = _seq(_0=_Conv(cat_size, out_size, 1, True))


class ResNet(nn.Module):
    """KGnet (KGnet.py:123-350): ResNet truncated after layer3 + decoder + keypoint heads + mask branch."""

    def __init__(self, block=None, layers=(3, 4, 6, 3), num_classes=1000, zero_init_residual=False, precision="fast"):
        """Same positional signature as the reference's ResNet(block, layers, num_classes, zero_init_residual)
        (KGnet.py:125); `num_classes` is unused there too (no fc layer is built).  `precision` is this package's
        only extra knob ("fast" | "exact" | "reference")."""
        super().__init__()
        if block is not None and not isinstance(block, type):     # ResNet([3, 4, 6, 3]) shorthand: first positional = layers
            block, layers = None, block
        if block is BasicBlock:
            _no_basic_block("ResNet(BasicBlock, ...)")
        self.blocks = tuple(int(b) for b in layers[:3])       # layer4 is never built (KGnet.py:131-137)
        self.conv1, self.bn1 = _Conv(3, 64, 7, False), _BN(64)
        inpl = 64
        for li, (planes, nb) in enumerate(zip((64, 128, 256), self.blocks)):
            blocks = []
            for b in range(nb):
                blocks.append(_Bottleneck(inpl, planes, downsample=(b == 0)))
                inpl = planes * 4
            setattr(self, f"layer{li + 1}", nn.ModuleList(blocks))
        self.c0_conv = _seq(_0=_Conv(3, 64, 3, True), _2=_Conv(64, 64, 3, True))
        self.skip_combine = nn.ModuleList([_Combination(64, 64, 128), _Combination(256, 64, 128),
                                           _Combination(512, 256, 512), _Combination(1024, 512, 1024)])
        self.seg_head = _seq(_0=_Conv(64, 64, 3, True), _2=_Conv(64, 1, 3, True))
        self.c4_up_conv = _seq(_0=_Conv(1024, 512, 3, True))
        self.c3_up_conv = _seq(_0=_Conv(512, 256, 3, True))
        self.c2_up_conv = _seq(_0=_Conv(256, 64, 3, True))
        self.c1_up_conv = _seq(_0=_Conv(64, 64, 3, True))
        self.c3_cat_refine = _seq(_0=_Conv(1024, 512, 1, True))
        self.c2_cat_refine = _seq(_0=_Conv(512, 256, 1, True))
        self.c1_cat_refine = _seq(_0=_Conv(128, 64, 1, True))
        self.c0_cat_refine = _seq(_0=_Conv(128, 64, 1, True))
        for s, c in ((3, 512), (2, 256), (1, 64), (0, 64)):
            for name, co in _HEADS:
                setattr(self, f"{name}_c{s}", _seq(_0=_Conv(c, c, 7, True), _2=_Conv(c, co, 7, True)))
        if zero_init_residual:                                 # KGnet.py:219-224
            for m in self.modules():
                if isinstance(m, _Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
        self.precision = precision
        self.export_feats = True
        self._fwd_serial = 0
        self._handle = None
        self._weights_sig = None
        self._flat_tensors = None
        self._ws = {}
        self._seg_ws = None
        self._last = None
        self.last_launches = 0

    # ---- weight upload --------------------------------------------------------------------------
    def _conv_list(self):
        """(state-dict prefix of the conv, prefix of the BatchNorm that follows it or None)."""
        out = [("conv1", "bn1"), ("c0_conv.0", None), ("c0_conv.2", None), ("seg_head.0", None), ("seg_head.2", None)]
        for li, nb in enumerate(self.blocks):
            for b in range(nb):
                p = f"layer{li + 1}.{b}"
                out += [(p + ".conv1", p + ".bn1"), (p + ".conv2", p + ".bn2"), (p + ".conv3", p + ".bn3")]
                if b == 0:
                    out.append((p + ".downsample.0", p + ".downsample.1"))
        for l in range(4):
            out += [(f"skip_combine.{l}.up.0", None), (f"skip_combine.{l}.cat_conv.0", None)]
        for n in ("c4_up_conv", "c3_up_conv", "c2_up_conv", "c1_up_conv", "c3_cat_refine", "c2_cat_refine", "c1_cat_refine",
                  "c0_cat_refine"):
            out.append((n + ".0", None))
        for s in range(4):
            for name, _ in _HEADS:
                out += [(f"{name}_c{s}.0", None), (f"{name}_c{s}.2", None)]
        return out

    def _signature(self):
        # (storage, version) of every parameter / buffer.  Walking the module tree costs 1.6 ms per call (measured), a tenth of a
        # bs4 step, so the (owner dict, key) slots are cached: load_state_dict / in-place edits bump _version, .to() / .cuda() change
        # data_ptr, and a Parameter object swapped by hand is seen too because the slot is looked up again on every call.
        if self._flat_tensors is None:
            self._flat_tensors = [(d, k) for m in self.modules() for d in (m._parameters, m._buffers) for k, v in d.items() if v is not None]
        return tuple((d[k].data_ptr(), d[k]._version) for d, k in self._flat_tensors)

    def _apply(self, fn, *args, **kw):
        self._flat_tensors = None
        return super()._apply(fn, *args, **kw)

    def _sync_weights(self):
        sig = self._signature()
        if self._handle is not None and sig == self._weights_sig:
            return
        L = _cabi.lib()
        if self._handle is None:
            h = C.c_void_p()
            blocks = (C.c_int * 3)(*self.blocks)
            _cabi.check(L.kg_net_create(C.byref(h), blocks))
            self._handle = h
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()
              if not k.endswith("num_batches_tracked")}
        fp = lambda t: t.numpy().ctypes.data_as(C.c_void_p) if t is not None else None
        for conv, bn in self._conv_list():
            w = sd[conv + ".weight"]
            b = sd.get(conv + ".bias")
            bnp = [sd[f"{bn}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")] if bn else [None] * 4
            co, ci, r, s = w.shape
            _cabi.check(L.kg_net_set_conv(self._handle, conv.encode(), fp(w), co, ci, r, s, fp(b), *[fp(t) for t in bnp], 1e-5))
        _cabi.check(L.kg_net_finalize(self._handle))
        self._weights_sig = sig
        self._last = None

    def __del__(self):
        try:
            if self._handle is not None:
                _cabi.lib().kg_net_destroy(self._handle)
        except Exception:
            pass

    def _precision_code(self):
        p = self.precision
        return PRECISIONS[p] if isinstance(p, str) else int(p)

    def _workspace(self, N, H, W, prec, device):
        key = (N, H, W, prec, str(device))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _cabi.lib().kg_net_workspace_bytes(self._handle, N, H, W, prec)
            if nbytes == 0:
                raise _cabi.KgError(-1, _cabi.lib().kg_last_error().decode())
            self._ws.clear()
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return ws

    # ---- KGnet.py:275-318 -----------------------------------------------------------------------
    def forward_dec(self, x):
        return self._forward_dec(x, False)

    def forward_dec_u8(self, images):
        """forward_dec of a uint8 [N,H,W,3] CUDA batch (cv2 BGR order) as it comes from the camera: the `x / 255 - 0.5` of test.py:92
        is folded into the two stem convs (exactly), no fp32 copy of the input is made.  Tensor-core precisions only."""
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
            raise ValueError(f"expected a uint8 [N,H,W,3] batch, got {tuple(images.shape)} {images.dtype}")
        if self._precision_code() == 0:
            raise RuntimeError("precision 'reference' runs on CUDA cores from the fp32 input: normalise with inference.preprocess_u8")
        return self._forward_dec(images, True)

    def _forward_dec(self, x, u8):
        if self.training:
            raise RuntimeError("kgnet_b200 implements the inference path: call .eval() first (train-mode BatchNorm is out of scope)")
        if not x.is_cuda:
            raise RuntimeError("kgnet_b200 needs CUDA tensors (no CPU fallback): move the model and input to cuda")
        if u8:
            x = x.detach().contiguous()
            N, H, W, c = x.shape
        else:
            x = x.detach().to(torch.float32).contiguous()
            N, c, H, W = x.shape
        assert c == 3, "input must be [N,3,H,W]"
        with torch.cuda.device(x.device):
            self._sync_weights()
            prec = self._precision_code()
            ws = self._workspace(N, H, W, prec, x.device)
            outs = []
            for s in range(4):
                hs, wsz = H >> s, W >> s
                outs.append([torch.empty(N, co, hs, wsz, dtype=torch.float32, device=x.device) for _, co in _HEADS])
            fsz = [(H, W), (H // 2, W // 2), (H // 4, W // 4), (H // 8, W // 8), (H // 16, W // 16)]
            feats = [torch.empty(N, cc, h, w, dtype=torch.float32, device=x.device) for cc, (h, w) in zip(_FEAT_C, fsz)] \
                if self.export_feats else None
            heads_p = (C.c_void_p * 12)(*[t.data_ptr() for o in outs for t in o])
            feats_p = (C.c_void_p * 5)(*[t.data_ptr() for t in feats]) if feats is not None else None
            nl = C.c_int(0)
            entry = _cabi.lib().kg_net_forward_dec_u8 if u8 else _cabi.lib().kg_net_forward_dec
            _cabi.check(entry(self._handle, x.data_ptr(), N, H, W, heads_p, feats_p, prec, ws.data_ptr(), ws.numel(),
                              torch.cuda.current_stream().cuda_stream, C.byref(nl)))
        self.last_launches = nl.value
        # every forward_dec gets a serial number: a _FeatList from an EARLIER pass must not match the workspace contents
        # of a later one (with export_feats=False the list holds no tensors that could tell the two apart)
        self._fwd_serial += 1
        feat_list = _FeatList(feats if feats is not None else [None] * 5)
        feat_list.owner_key = (id(self), N, H, W, prec, self._fwd_serial)
        self._last = (feat_list.owner_key, ws, [None if t is None else (t.data_ptr(), t._version) for t in feat_list])
        return outs[0], outs[1], outs[2], outs[3], feat_list

    # ---- KGnet.py:246-256 -----------------------------------------------------------------------
    def get_patches(self, box, feat):
        """Crop of one [C,h,w] feature map for a box given in normalised (y1,x1,y2,x2) coordinates, or None when the
        crop is thinner than 2 px.  Same rounding as the device path (`patch_rect` in csrc/net.cu): half-to-even on
        float32 products, clamped to the map."""
        rect = patch_rect(box, feat.shape[-2], feat.shape[-1])
        if rect is None:
            return None
        top, left, bottom, right = rect
        return feat[None, :, top:bottom, left:right]

    # ---- KGnet.py:321-350 -----------------------------------------------------------------------
    def forward_seg(self, feat_seg, bboxes):
        """feat_seg: the list returned by forward_dec (any list of 5 fp32 NCHW CUDA tensors works too);
        bboxes: per image an (M,5) array [y1,x1,y2,x2,score] or an empty list.
        Returns [mask_patches, mask_dets] with the reference's nesting."""
        return self.forward_seg_packed(feat_seg, bboxes).as_lists()

    def forward_seg_packed(self, feat_seg, bboxes):
        """Same computation as forward_seg, but returns a SegResult: all mask patches in ONE device buffer plus their
        geometry, without creating a Python tensor object per box (the reference-style nested lists are built on demand
        by SegResult.as_lists())."""
        nimg = len(bboxes)
        L = _cabi.lib()
        internal = (isinstance(feat_seg, _FeatList) and self._last is not None and feat_seg.owner_key == self._last[0] and
                    all((t is None and s is None) or (t is not None and s is not None and (t.data_ptr(), t._version) == s)
                        for t, s in zip(feat_seg, self._last[2])))
        if internal:
            _, N, H, W, prec, _ = self._last[0]
            ws = self._last[1]
            device = ws.device
        else:
            f0 = feat_seg[0]
            if f0 is None or not f0.is_cuda:
                raise RuntimeError("forward_seg needs CUDA feature tensors (or the list returned by forward_dec)")
            N, _, H, W = f0.shape
            device = f0.device
            with torch.cuda.device(device):
                self._sync_weights()
                prec = self._precision_code()
                ws = self._workspace(N, H, W, prec, device)
                fl = [t.detach().to(torch.float32).contiguous() for t in feat_seg]
                fp = (C.c_void_p * 5)(*[t.data_ptr() for t in fl])
                _cabi.check(L.kg_net_import_feats(self._handle, fp, N, H, W, prec, ws.data_ptr(), ws.numel(),
                                                  torch.cuda.current_stream().cuda_stream))
            self._last = None
        if nimg > N:
            raise ValueError(f"{nimg} box lists for a batch of {N} images")
        counts = np.zeros(N, np.int32)
        rows = []
        for i in range(nimg):
            if len(bboxes[i]) == 0:
                continue
            b = np.asarray(bboxes[i], np.float64).reshape(-1, 5)
            counts[i] = len(b)
            rows.append(b)
        total = int(counts.sum())
        if total == 0:
            self.last_launches = 0
            self.last_masks = torch.empty(0, dtype=torch.float32, device=device)
            return SegResult(nimg, counts, np.zeros((0, 5)), self.last_masks, np.zeros(0, np.int32), np.zeros((0, 2), np.int32),
                             np.zeros(0, np.int64), np.zeros(0, np.int32))
        boxes = np.ascontiguousarray(np.concatenate(rows, 0))
        seg_bytes, mask_floats, n_masks = C.c_size_t(0), C.c_longlong(0), C.c_int(0)
        mask_index = np.zeros(total, np.int32); mask_hw = np.zeros((total, 2), np.int32); mask_off = np.zeros(total, np.int64)
        mask_pitch = np.zeros(total, np.int32)
        with torch.cuda.device(device):
            _cabi.check(L.kg_net_seg_prepare(self._handle, N, H, W, counts.ctypes.data, boxes.ctypes.data, C.byref(seg_bytes),
                                             C.byref(mask_floats), C.byref(n_masks), mask_index.ctypes.data, mask_hw.ctypes.data,
                                             mask_off.ctypes.data, mask_pitch.ctypes.data))
            if self._seg_ws is None or self._seg_ws.numel() < seg_bytes.value or self._seg_ws.device != device:
                self._seg_ws = torch.empty(int(seg_bytes.value * 1.25) + 1024, dtype=torch.uint8, device=device)
            masks = torch.empty(max(1, mask_floats.value), dtype=torch.float32, device=device)
            nl = C.c_int(0)
            if getattr(self, "_seg_launch_event", None) is not None:
                self._seg_launch_event.record()
            _cabi.check(L.kg_net_forward_seg(self._handle, ws.data_ptr(), self._seg_ws.data_ptr(), self._seg_ws.numel(),
                                             masks.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(nl)))
        self.last_launches = nl.value
        self.last_masks = masks
        return SegResult(nimg, counts, boxes, masks, mask_index, mask_hw, mask_off, mask_pitch)

    # ---- KGnet.py:269-272 -----------------------------------------------------------------------
    def forward(self, x, bboxes):
        dec0, dec1, dec2, dec3, feat_seg = self.forward_dec(x)
        seg = self.forward_seg(feat_seg, bboxes)
        return dec0, dec1, dec2, dec3, seg


def patch_rect(box_norm, h, w):
    """Integer crop rectangle (top, left, bottom, right) of a normalised box on an h x w map (KGnet.py:246-256), or
    None.  float32 arithmetic, round-half-to-even, like NumPy on the reference's float32 scalars."""
    b = np.asarray(box_norm, np.float32)
    lo = np.rint(b[:2] * np.float32([h, w])).astype(np.int64)
    hi = np.rint(b[2:4] * np.float32([h, w])).astype(np.int64)
    lo = np.maximum(lo, 0)
    hi = np.minimum(hi, [h - 1, w - 1])
    if (hi - lo < 2).any():
        return None
    return int(lo[0]), int(lo[1]), int(hi[0]), int(hi[1])


class SegResult:
    """Packed output of forward_seg: `masks` is one fp32 device buffer; patch k of the concatenated box list lives at
    masks[off[slot] + y * pitch[slot] + x] for slot = index[k] >= 0 (index -1: the box was skipped, KGnet.py:341-342)."""

    def __init__(self, nimg, counts, boxes, masks, index, hw, off, pitch):
        self.nimg, self.counts, self.boxes, self.masks = nimg, counts, boxes, masks
        self.index, self.hw, self.off, self.pitch = index, hw, off, pitch

    def patch(self, k):
        slot = int(self.index[k])
        if slot < 0:
            return None
        h, w = int(self.hw[slot, 0]), int(self.hw[slot, 1])
        return torch.as_strided(self.masks, (h, w), (int(self.pitch[slot]), 1), int(self.off[slot]))

    def as_lists(self):
        """[mask_patches, mask_dets] nested like the reference's forward_seg return value (KGnet.py:346-350)."""
        mask_patches = [[] for _ in range(self.nimg)]
        mask_dets = [[] for _ in range(self.nimg)]
        out = _SegLists([mask_patches, mask_dets])
        out.packed = self
        k = 0
        for i in range(self.nimg):
            for _ in range(int(self.counts[i])):
                pt = self.patch(k)
                if pt is not None:
                    mask_patches[i].append(pt)
                    mask_dets[i].append(torch.Tensor(np.append(self.boxes[k, :4], self.boxes[k, 4])))
                k += 1
        return out

    def paste_geometry(self):
        """(buffer, float offsets, row pitches, (h, w)) of the kept patches in as_lists() order (for kg_paste_masks)."""
        slots = self.index[self.index >= 0]
        return self.masks, self.off[slots], self.pitch[slots], self.hw[slots]


class _SegLists(list):
    """[mask_patches, mask_dets] (a plain list to every caller) that remembers the packed buffer its patches view."""
    packed = None


class _FeatList(list):
    """[c0..c4] as returned by forward_dec; remembers which forward pass produced it so that forward_seg can
    read the features from the device workspace in their internal layout instead of re-importing them."""
    owner_key = None


def _no_basic_block(name):
    raise NotImplementedError(f"{name}: the reference's decoder hard-codes Bottleneck widths (c2/c3/c4 = 256/512/1024, "
                              "KGnet.py:150-158), so its BasicBlock variants cannot run forward_dec either")


def resnet18(pretrained=False, **kwargs):
    _no_basic_block("resnet18")


def resnet34(pretrained=False, **kwargs):
    _no_basic_block("resnet34")


class Bottleneck:
    """Block-type marker with the reference's name and expansion (KGnet.py:64-66) for `ResNet(Bottleneck, layers)`."""
    expansion = 4


class BasicBlock:
    """Marker only: the reference's BasicBlock variants cannot run its own decoder (see _no_basic_block)."""
    expansion = 1


model_urls = {   # KGnet.py:12-18 (never fetched here: see _load_pretrained)
    "resnet50": "https://download.pytorch.org/models/resnet50-19c8e357.pth",
    "resnet101": "https://download.pytorch.org/models/resnet101-5d3b4d8f.pth",
    "resnet152": "https://download.pytorch.org/models/resnet152-b121ed2d.pth",
}


def _load_pretrained(model, arch):
    """`pretrained=True` in the reference pulls the ImageNet trunk with model_zoo.load_url and loads it with strict=False
    (KGnet.py:383-386); test.py:53 / eval.py:31 construct the model that way and then overwrite EVERY tensor with
    `load_weights(end_model.pth)`.  This package never touches the network: a local copy of the torchvision checkpoint
    is used when one exists ($KGNET_PRETRAINED, or the torch hub cache), otherwise construction proceeds with the random
    initialisation and a warning -- the result after load_weights() is identical."""
    import os
    import warnings
    name = os.path.basename(model_urls[arch])
    cands = [os.environ.get("KGNET_PRETRAINED", ""),
             os.path.join(torch.hub.get_dir(), "checkpoints", name),
             os.path.join(os.path.expanduser("~"), ".torch", "models", name)]
    for c in cands:
        if c and os.path.isfile(c):
            model.load_state_dict(torch.load(c, map_location="cpu"), strict=False)
            return True
    warnings.warn(f"KGnet.{arch}(pretrained=True): no local copy of {name} (offline; set KGNET_PRETRAINED=<file>): keeping the "
                  "random initialisation. Load a trained checkpoint with load_state_dict().", stacklevel=3)
    return False


def _make(arch, layers, pretrained, kwargs):
    model = ResNet(Bottleneck, layers, **kwargs)
    if pretrained:
        _load_pretrained(model, arch)
    return model


def resnet50(pretrained=False, **kwargs):
    """KGnet.py:377-386."""
    return _make("resnet50", (3, 4, 6, 3), pretrained, kwargs)


def resnet101(pretrained=False, **kwargs):
    """KGnet.py:389-398."""
    return _make("resnet101", (3, 4, 23, 3), pretrained, kwargs)


def resnet152(pretrained=False, **kwargs):
    """KGnet.py:401-410."""
    return _make("resnet152", (3, 8, 36, 3), pretrained, kwargs)
