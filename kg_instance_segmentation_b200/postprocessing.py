"""Drop-in for the reference's postprocessing.py: same function names, argument meaning and return
types; every function runs the CUDA decode kernels (csrc/decode.cu) through the C-ABI.

`decode_batched` is the new batched entry point (the reference decodes batch element 0 only,
postprocessing.py:138-140); `get_skeletons_and_masks` keeps the reference's single-image contract.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from . import config as cfg


MAX_CAPACITY = 8192      # largest per-(image, scale) peak list / per-image box list the device kernels hold in shared memory


def _as_cuda_f32(t, device=None):
    """Contiguous fp32 CUDA tensor: CUDA inputs stay on THEIR device, host inputs go to `device` (default: current)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    if not torch.cuda.is_available():
        raise RuntimeError("kg_instance_segmentation_b200 needs a CUDA device (no CPU fallback)")
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


@dataclass
class DecodeResult:
    """Device-resident outputs of one decode call (torch tensors own the memory)."""
    dets: torch.Tensor          # [N, max_boxes, 5] f64
    det_count: torch.Tensor     # [N] i32
    boxes: Optional[torch.Tensor] = None
    box_count: Optional[torch.Tensor] = None
    skeletons: Optional[torch.Tensor] = None     # [N, S, max_peaks, 5, 3] f64
    skel_count: Optional[torch.Tensor] = None    # [N, S]
    skel_keep: Optional[torch.Tensor] = None     # [N, S, max_peaks] u8
    peak_conf: Optional[torch.Tensor] = None
    peak_key: Optional[torch.Tensor] = None
    peak_count: Optional[torch.Tensor] = None
    heat: Optional[list] = None
    vote: Optional[list] = None
    status: Optional[torch.Tensor] = None
    packed: Optional[torch.Tensor] = None        # [N, packed_k + 1, 5] f64: all-gather record (row 0 = count, rows stored)
    n_launches: int = 0
    meta: Optional[torch.Tensor] = None          # [N + 1] i32 = det_count | status (one buffer: one D2H fetches both)
    h_meta: Optional[torch.Tensor] = None        # pinned host copies
    h_dets: Optional[torch.Tensor] = None
    _guess: int = 64                             # detection rows per image fetched speculatively together with the counts
    _fetch: Optional[tuple] = None               # (event, rows) of the copies enqueued by prefetch()
    _host: Optional[tuple] = None                # (status, counts, rows) once the copies have landed

    def invalidate(self):
        """A new decode has been enqueued into these buffers."""
        self._fetch = self._host = None

    def prefetch(self):
        """Enqueue, on the current stream (i.e. right behind the decode), the D2H copies the host will need: status + counts and
        the first rows of every image's detection list, into pinned memory; record an event.  No synchronisation: a pipelined
        caller enqueues the next batch before it waits."""
        if self._fetch is not None or self._host is not None:
            return
        dev = self.dets.device
        N, cap = self.dets.shape[0], self.dets.shape[1]
        if self.h_meta is None:
            self.h_meta = torch.empty(N + 1, dtype=torch.int32).pin_memory()
            self.h_dets = torch.empty(N * cap * 5, dtype=torch.float64).pin_memory()
        g = max(1, min(self._guess, cap))
        with torch.cuda.device(dev):
            self.h_meta.copy_(self.meta, non_blocking=True)
            self.h_dets[:N * g * 5].view(N, g, 5).copy_(self.dets[:, :g], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
        self._fetch = (ev, g)

    def _landed(self):
        """(status, counts [N] i32, rows [N, m, 5] f64) on the host: ONE blocking wait per decode (a second copy only when an image
        holds more detections than the speculative fetch covered)."""
        if self._host is None:
            self.prefetch()
            ev, g = self._fetch
            ev.synchronize()
            N, cap = self.dets.shape[0], self.dets.shape[1]
            cnt = self.h_meta[:N].numpy().copy()
            st = int(self.h_meta[N])
            m = int(cnt.max()) if N else 0
            rows = self.h_dets[:N * g * 5].view(N, g, 5)
            if m > g:
                rows = self.h_dets[:N * m * 5].view(N, m, 5)
                rows.copy_(self.dets[:, :m])             # blocking copy on the current stream
            self._guess = int(min(cap, max(64, 1 << int(np.ceil(np.log2(max(1, m) * 1.25))))))
            self._host = (st, cnt, rows.numpy())
            self._fetch = None
        return self._host

    def overflow(self) -> int:
        """0, or the status bits of a list overflow (bit0: a peak list, bit1: a box list).  Synchronises."""
        if self.meta is None:
            return int(self.status.item())
        return self._landed()[0]

    def check(self):
        st = self.overflow()
        if st & 3:
            raise _cabi.KgError(-3, f"decode list overflow (status={st}: bit0 peaks, bit1 boxes); raise max_peaks/max_boxes")

    def detections(self) -> List[Optional[np.ndarray]]:
        """Per image: (M,5) float64 array in NMS keep order, or None (nms.py:8-9) when there are no boxes."""
        self.check()
        if self.meta is None:
            cnt = self.det_count.cpu().numpy()
            m = int(cnt.max()) if len(cnt) else 0
            d = self.dets[:, :max(m, 1)].cpu().numpy()
        else:
            _, cnt, d = self._landed()
        return [d[i, :c].copy() if c > 0 else None for i, c in enumerate(cnt)]


class Decoder:
    """Holds the workspace / output buffers for one (N, scale shapes) configuration so that repeated
    calls launch kernels only (no allocation, no host sync)."""

    def __init__(self, N, shapes: Sequence[tuple], box_scales: Sequence[int] = None, max_peaks=4096, max_boxes=4096,
                 nms_thresh=0.5, peak_thresh=cfg.PEAK_THRESH, debug=False, device="cuda", packed_k=0):
        self.L = _cabi.lib()
        self.N, self.shapes = int(N), [tuple(map(int, s)) for s in shapes]
        S = len(self.shapes)
        self.box_scales = list(box_scales) if box_scales is not None else list(cfg.BOX_SCALES[:S])
        self.cfg = _cabi.DecodeConfig(self.N, S, int(max_peaks), int(max_boxes), float(nms_thresh), float(peak_thresh))
        self.sc = (_cabi.DecodeScale * S)()
        for s, (h, w) in enumerate(self.shapes):
            self.sc[s].H, self.sc[s].W, self.sc[s].box_scale = h, w, int(self.box_scales[s])
        ws = self.L.kg_decode_workspace_bytes(C.byref(self.cfg), self.sc)
        if ws == 0:
            raise _cabi.KgError(-1, self.L.kg_last_error().decode())
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.max_peaks, self.max_boxes = int(max_peaks), int(max_boxes)
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=dev)
        self.debug = debug
        f64, i32 = torch.float64, torch.int32
        meta = torch.zeros(self.N + 1, dtype=i32, device=dev)
        r = DecodeResult(dets=torch.zeros(self.N, max_boxes, 5, dtype=f64, device=dev), det_count=meta[:self.N], status=meta[self.N:],
                         meta=meta)
        if debug:
            r.boxes = torch.zeros(self.N, max_boxes, 5, dtype=f64, device=dev)
            r.box_count = torch.zeros(self.N, dtype=i32, device=dev)
            r.peak_conf = torch.zeros(self.N, S, max_peaks, dtype=f64, device=dev)
            r.peak_key = torch.zeros(self.N, S, max_peaks, dtype=i32, device=dev)
            r.peak_count = torch.zeros(self.N, S, dtype=i32, device=dev)
            r.skel_keep = torch.zeros(self.N, S, max_peaks, dtype=torch.uint8, device=dev)
            r.heat = [torch.zeros(self.N, 5, h, w, dtype=f64, device=dev) for h, w in self.shapes]
            r.vote = [torch.zeros(self.N, 5, h, w, dtype=f64, device=dev) for h, w in self.shapes]
        if packed_k > 0:
            r.packed = torch.zeros(self.N, packed_k + 1, 5, dtype=f64, device=dev)
        r.skeletons = torch.zeros(self.N, S, max_peaks, 5, 3, dtype=f64, device=dev)
        r.skel_count = torch.zeros(self.N, S, dtype=i32, device=dev)
        self.result = r
        o = _cabi.DecodeOutputs()
        p = lambda t: t.data_ptr() if t is not None else None
        o.d_dets, o.d_det_count, o.d_boxes, o.d_box_count = p(r.dets), p(r.det_count), p(r.boxes), p(r.box_count)
        o.d_skeletons, o.d_skel_count, o.d_skel_keep = p(r.skeletons), p(r.skel_count), p(r.skel_keep)
        o.d_peak_conf, o.d_peak_key, o.d_peak_count = p(r.peak_conf), p(r.peak_key), p(r.peak_count)
        for s in range(S):
            o.d_heat[s] = p(r.heat[s]) if debug else None
            o.d_vote[s] = p(r.vote[s]) if debug else None
        o.d_status = p(r.status)
        o.d_det_packed, o.det_packed_k = p(r.packed), int(packed_k)
        self.out = o

    def __call__(self, heads, stream: Optional[torch.cuda.Stream] = None) -> DecodeResult:
        """heads: per scale (kp [N,5,H,W], short [N,10,H,W], mid [N,40,H,W]) fp32 contiguous CUDA tensors.
        Enqueues on `stream` (default: torch's current stream); does not synchronise."""
        keep = []
        for s, (kp, sh, mid) in enumerate(heads):
            h, w = self.shapes[s]
            for t, c in ((kp, 5), (sh, 10), (mid, 40)):
                if (t.dtype != torch.float32 or t.device != self.device or not t.is_contiguous() or
                        tuple(t.shape) != (self.N, c, h, w)):
                    raise ValueError(f"scale {s}: expected contiguous float32 [{self.N},{c},{h},{w}] on {self.device}, got "
                                     f"{tuple(t.shape)} {t.dtype} {t.device}")
            self.sc[s].d_kp, self.sc[s].d_short, self.sc[s].d_mid = kp.data_ptr(), sh.data_ptr(), mid.data_ptr()
            keep.append((kp, sh, mid))
        nl = C.c_int(0)
        with torch.cuda.device(self.device):           # kernels must launch on the device that owns the buffers
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            _cabi.check(self.L.kg_decode(C.byref(self.cfg), self.sc, C.byref(self.out), self.workspace.data_ptr(),
                                         self.workspace.numel(), st.cuda_stream, C.byref(nl)))
        self.result.n_launches = nl.value
        self.result.invalidate()
        self._keepalive = keep
        return self.result


_decoders = {}


def _decoder(N, shapes, box_scales, device, **kw) -> Decoder:
    key = (N, tuple(shapes), tuple(box_scales), str(device), tuple(sorted(kw.items())))
    d = _decoders.get(key)
    if d is None:
        if len(_decoders) > 8:
            _decoders.clear()
        d = _decoders[key] = Decoder(N, shapes, box_scales, device=device, **kw)
    return d


def run_with_growth(make_decoder, heads, max_peaks, max_boxes):
    """Run a decode; when a bounded device list overflowed, re-run with doubled capacities (the reference has no caps:
    postprocessing.py builds Python lists).  Raises only past MAX_CAPACITY entries per (image, scale) list."""
    while True:
        res = make_decoder(max_peaks, max_boxes)(heads)
        st = res.overflow() & 3
        if not st:
            return res
        grown = False
        if st & 1 and max_peaks < MAX_CAPACITY:
            max_peaks, grown = max_peaks * 2, True
        if st & 2 and max_boxes < MAX_CAPACITY:
            max_boxes, grown = max_boxes * 2, True
        if not grown:
            res.check()


def decode_batched(heads, nms_thresh=0.5, max_peaks=4096, max_boxes=4096, debug=False) -> DecodeResult:
    """test.py:105-116 for a whole batch: heads = [[kp_s, short_s, mid_s] for s in 0..3] (NCHW torch tensors).
    List capacities grow on overflow (up to 8192 peaks per image-scale)."""
    first = heads[0][0]
    dev = first.device if isinstance(first, torch.Tensor) and first.is_cuda else None
    hs = [tuple(_as_cuda_f32(t, dev) for t in h) for h in heads]
    dev = hs[0][0].device
    N = hs[0][0].shape[0]
    shapes = [tuple(h[0].shape[2:]) for h in hs]
    mk = lambda mp, mb: _decoder(N, shapes, cfg.BOX_SCALES[:len(hs)], dev, nms_thresh=float(nms_thresh), max_peaks=mp,
                                 max_boxes=mb, debug=debug)
    return run_with_growth(mk, hs, max_peaks, max_boxes)


def _skeleton_list(res: DecodeResult, n, s):
    cnt = int(res.skel_count[n, s].item())
    sk = res.skeletons[n, s, :cnt].cpu().numpy()
    return [sk[i].copy() for i in range(cnt)]


def get_skeletons_and_masks(kp_maps, short_offsets, mid_offsets):
    """postprocessing.py:129-147: batch x {5,10,40} x H x W tensors -> list of (5,3) float64 [x,y,conf]
    skeletons of batch element 0 (missing keypoints are all-zero rows)."""
    dev = kp_maps.device if isinstance(kp_maps, torch.Tensor) and kp_maps.is_cuda else None
    kp, sh, mid = (_as_cuda_f32(t, dev)[0:1].contiguous() for t in (kp_maps, short_offsets, mid_offsets))
    mk = lambda mp, mb: _decoder(1, [tuple(kp.shape[2:])], (1,), kp.device, max_peaks=mp, max_boxes=mb, debug=False)
    res = run_with_growth(mk, [(kp, sh, mid)], 4096, 4096)
    return _skeleton_list(res, 0, 0)


def _boxes_host(skeletons, scale, apply_refine):
    n = len(skeletons)
    if n == 0:
        return np.zeros((0,), np.uint8), np.zeros((0, 5), np.float64)
    sk = np.ascontiguousarray(np.asarray(skeletons, np.float64).reshape(n, 5, 3))
    keep = np.zeros(n, np.uint8)
    boxes = np.zeros((n, 5), np.float64)
    nb = C.c_int(0)
    _cabi.check(_cabi.lib().kg_skeletons_to_boxes_host(sk.ctypes.data, n, int(scale), int(apply_refine), keep.ctypes.data,
                                                       boxes.ctypes.data, C.byref(nb)))
    return keep, boxes[:nb.value]


def refine_skeleton(skeletons):
    """postprocessing.py:150-159."""
    keep, _ = _boxes_host(skeletons, 1, False)
    return [s for s, k in zip(skeletons, keep) if k]


def skeleton_to_box(skeletons, scale):
    """postprocessing.py:164-242.  Like the reference this scales the skeletons' xy IN PLACE."""
    _, boxes = _boxes_host(skeletons, scale, False)
    for s in skeletons:
        s[:, :2] *= scale
    return [list(b) for b in boxes]


def gather_skeleton_single(skeleton0, skeleton1, skeleton2, skeleton3):
    """postprocessing.py:245-252."""
    return tuple(np.asarray(skeleton_to_box(s, sc)) for s, sc in zip((skeleton0, skeleton1, skeleton2, skeleton3),
                                                                     cfg.BOX_SCALES))


def gather_skeleton(skeleton0, skeleton1, skeleton2, skeleton3):
    """postprocessing.py:255-261."""
    b = []
    for s, sc in zip((skeleton0, skeleton1, skeleton2, skeleton3), cfg.BOX_SCALES):
        b += skeleton_to_box(s, sc)
    return np.asarray(b)
