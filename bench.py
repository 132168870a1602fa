#!/usr/bin/env python
"""Benchmark of the KGnet inference hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pipeline|decode] [--impl reference]

One "step" = one pass of the hot path over one batch (bs 32 per GPU, 512x512 synthetic).  Prints ONE JSON line.
Multi-GPU: launched under torchrun, one rank per GPU; images are sharded (weak scaling: bs 32 per rank), the only
collective is the all-gather of the padded detection list.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "images/s"
# BASELINE.json configs: cfg2 (= cfg3 per GPU) is the configuration the metric is quoted on; cfg4 stresses top-K + grouping
CONFIGS = {
    "cfg2": dict(bs=32, hw=512, cells=40, side=(24, 110), gap=12,
                 metric="images/sec at 512x512 bs32 (KGnet inference hot path)"),
    "cfg1": dict(bs=1, hw=512, cells=40, side=(24, 110), gap=12,
                 metric="images/sec at 512x512 bs1 (KGnet inference hot path, BASELINE config 1: the reference's own test.py case)"),
    "cfg4": dict(bs=8, hw=1024, cells=500, side=(16, 40), gap=6,
                 metric="images/sec at 1024x1024 bs8, ~500 cells/img (KGnet inference hot path, BASELINE config 4)"),
}
METRIC = CONFIGS["cfg2"]["metric"]
BS, HW_IN, CELLS, SIDE, GAP = 32, 512, 40, (24, 110), 12
MAX_PEAKS, MAX_BOXES, MAX_DETS = 4096, 4096, 1024


def set_config(name, world=1, scaling="weak"):
    """Selects the workload; strong scaling divides the global batch over the ranks."""
    global METRIC, BS, HW_IN, CELLS, SIDE, GAP
    c = CONFIGS[name]
    METRIC, BS, HW_IN, CELLS, SIDE, GAP = c["metric"], c["bs"], c["hw"], c["cells"], c["side"], c["gap"]
    if scaling == "strong":
        if BS % world:
            raise SystemExit(f"strong scaling: global batch {BS} is not divisible by {world} ranks")
        BS //= world
# algorithmic HBM bytes per pixel per scale of the decode: vote reads 5+10 f32 and writes 5 x 8 B
# accumulators (100 B), blur+peak reads the accumulators back once (40 B)
VOTE_BYTES_PX, BLUR_BYTES_PX = 100, 40


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def profiled_traffic(key=None):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary (profiles/*_traffic.json);
    None when no capture has been committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    state = d.get("code_state")
    if key is not None:
        d = d.get(key)
        if d is None:
            return None, None
    # the capture is a separate ncu run of the same command (a number taken under a profiler is never a bench value): the source
    # says which code state and how many launches it covers, so that a stale capture is visible in the line
    src = os.path.basename(files[-1]) + (f" ({d.get('launches')} launches" + (f", code {state}" if state else "") + ")")
    return d.get("dram_bytes_per_launch"), src


def planted_batch(n_distinct=4):
    """Teacher-forced decode load: planted 40-cell scenes, bs 32 built from n_distinct scenes."""
    from kg_instance_segmentation_b200 import synthetic
    base = [synthetic.planted_scene(100 + i, HW_IN, HW_IN, CELLS, side=SIDE, gap=GAP)[0] for i in range(n_distinct)]
    scenes = [base[i % n_distinct] for i in range(BS)]
    return base, [tuple(np.stack([sc[s][k] for sc in scenes]) for k in range(3)) for s in range(4)]


# ---------------------------------------------------------------------------------------------------------
def run_decode(args, rank, world, dist):
    import torch
    from kg_instance_segmentation_b200 import _cabi, postprocessing
    dev = torch.device("cuda", torch.cuda.current_device())
    base, host = planted_batch()
    host_pinned = [tuple(torch.from_numpy(a).pin_memory() for a in h) for h in host]
    dev_heads = [tuple(t.to(dev) for t in h) for h in host_pinned]
    shapes = [tuple(h[0].shape[2:]) for h in dev_heads]
    dec = postprocessing.Decoder(BS, shapes, max_peaks=MAX_PEAKS, max_boxes=MAX_BOXES, packed_k=MAX_DETS)
    stage_in = [tuple(torch.empty_like(t) for t in h) for h in dev_heads]
    out_host = torch.empty(BS, MAX_DETS + 1, 5, dtype=torch.float64).pin_memory()
    gathered = torch.empty(world * BS, MAX_DETS + 1, 5, dtype=torch.float64, device=dev) if world > 1 else None

    def step_device():
        r = dec(dev_heads)
        if world > 1:   # the single collective of the path: ONE all-gather of the fixed-size detection records (row 0 = count)
            dist.all_gather_into_tensor(gathered, r.packed)
        return r

    def step_e2e():
        for hs, ds in zip(host_pinned, stage_in):
            for h, d in zip(hs, ds):
                d.copy_(h, non_blocking=True)
        r = dec(stage_in)
        if world > 1:
            dist.all_gather_into_tensor(gathered, r.packed)
        out_host.copy_(r.packed, non_blocking=True)
        return r

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    for _ in range(max(args.warmup, 3)):
        r = step_device()
    r.check()
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    ms, r = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else None
    r.check()
    n_det = int(r.det_count.sum().item())
    launches = r.n_launches * args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    h2d = sum(t.numel() * 4 for h in host_pinned for t in h)
    d2h = out_host.numel() * 8

    # per-kernel time of the dominant kernel, live, with CUDA events on the launching stream
    _cabi.timing_enable(True)
    for _ in range(args.steps):
        step_device()
    st_ms, st_cnt = _cabi.timing_collect()
    _cabi.timing_enable(False)
    px = BS * sum(h * w for h, w in shapes)
    # stage 0 is ONE timed region (vote + prefilter of the four scales: 8 launches back to back)
    names = {0: "vote_kernel + blur32_candidates_kernel (4 scales)", 1: "exact_peaks_kernel", 2: "group_kernel", 3: "nms_kernel"}
    stage = {names[i]: {"ms_per_step": float(st_ms[i]) / args.steps, "launches_per_step": int(st_cnt[i]) // args.steps} for i in names}
    # the HBM-bound part of the decode is heat-map -> peak list (140 B per pixel and scale: vote reads 15 f32 and
    # writes 5 x 8 B accumulators, blur+peak reads them back once); grouping / NMS are latency-bound list kernels (~0 bytes)
    alg_bytes = px * (VOTE_BYTES_PX + BLUR_BYTES_PX)
    pk, pk_kind = peaks()
    dom_ms = float(st_ms[0] + st_ms[1]) / args.steps
    dom_launches = 2 * len(shapes) + 1          # a vote and a prefilter launch per scale + the candidate re-evaluation
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_src = profiled_traffic("decode")
    roofline = {"kernel": "vote_kernel + blur32_candidates_kernel + exact_peaks_kernel (head maps -> peak lists)", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": pk_kind + " (burst copy)", "ms_per_step": round(dom_ms, 4), "launches_per_step": dom_launches,
                "algorithmic_bytes_per_step": alg_bytes, "algorithmic_bytes_per_launch": alg_bytes // max(1, dom_launches),
                "stages": stage}
    out = {
        "metric": METRIC, "value": round(world * BS * args.steps / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": f"synthetic (planted {CELLS}-cell scenes, teacher-forced head maps)",
        "config": {"workload": f"decode-only: bs{BS}/GPU {HW_IN}x{HW_IN} head maps (4 scales) -> vote+blur+peak+group+boxes+NMS; inputs 2.45 GB > L2, no flush needed",
                   "config": args.config, "global_batch": world * BS, "detections_per_step": n_det},
        "e2e": {"value": round(world * BS * args.steps / (ms_e2e * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_decode(base)
    return out


# ---------------------------------------------------------------------------------------------------------
# FLOPs of forward_dec per image at 512x512 come from the plan census (kg_net_plan_info); SURVEY.md §8d: 1357.31 GF.
STAGE_NAMES = {0: "vote_blur_prefilter", 1: "exact_peaks", 2: "group", 3: "nms", 8: "conv_cuda_core", 9: "tc_backbone", 10: "tc_decoder",
               11: "tc_heads_l1", 12: "tc_heads_l2", 13: "bilinear", 14: "maxpool", 15: "export_feats", 16: "forward_seg"}


def seg_flops(dets, H, W):
    """Algorithmic FLOPs (2*MACs, true crop sizes) of forward_seg for a list of per-image detections: KGnet.py:258-267,321-350."""
    from kg_instance_segmentation_b200.KGnet import patch_rect
    up_in, lvl_out, feat_c = (64, 256, 512, 1024), (64, 64, 256, 512), (64, 64, 256, 512, 1024)
    total, boxes = 0.0, 0
    for d in dets:
        if d is None:
            continue
        for row in d:
            b = np.asarray(row[:4], np.float32) / np.float32([H, W, H, W])
            areas = []
            for l in range(5):
                r = patch_rect(b, H >> l, W >> l)
                if r is None:
                    break
                areas.append((r[2] - r[0]) * (r[3] - r[1]))
            if not areas:
                continue
            boxes += 1
            for l in range(len(areas) - 1):     # level l receives the upsampled level l+1: 3x3 up conv + 1x1 conv over the concat
                total += 2.0 * areas[l] * (up_in[l] * lvl_out[l] * 9 + (feat_c[l] + lvl_out[l]) * lvl_out[l])
            total += 2.0 * areas[0] * (64 * 64 * 9 + 64 * 9)      # seg_head
    return total, boxes


def gpu_library_baseline(dev, steps=3):
    """forward_dec of the SAME graph through PyTorch's library kernels (cuDNN convs, ATen pooling / resize) on the same GPU:
    the implementation the hand-written kernels have to beat.  fp32 with TF32 off (the reference's accuracy),
    TF32 (PyTorch's default for convs) and fp16 autocast + channels_last (the fastest library path)."""
    import torch
    from oracle import kg_oracle as O
    sd = {k: v.to(dev) for k, v in O.make_state_dict(seed=0).items()}
    torch.manual_seed(0)
    x = torch.rand(BS, 3, HW_IN, HW_IN, device=dev) - 0.5
    out = {}
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, half in (("fp32_tf32_off", False, False), ("tf32", True, False), ("fp16_autocast_channels_last", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            sdv, xv = sd, x
            if half:
                sdv = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
                xv = x.contiguous(memory_format=torch.channels_last)
            def run():
                if half:
                    with torch.autocast("cuda", dtype=torch.float16):
                        return O.forward_dec(sdv, xv)
                return O.forward_dec(sdv, xv)
            try:
                for _ in range(2):
                    r = run()
                del r
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    r = run()
                    del r
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"ms_per_step": round(ms, 3), "images_per_s": round(BS / (ms * 1e-3), 1)}
            except Exception as ex:       # e.g. out of memory on a small part: report, do not fail the bench
                out[name] = {"error": str(ex)[:200]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = prev
    out["what"] = f"forward_dec only, bs{BS} {HW_IN}x{HW_IN}, torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}"
    return out


def run_pipeline(args, rank, world, dist):
    import ctypes as C
    import torch
    from kg_instance_segmentation_b200 import _cabi, synthetic
    from kg_instance_segmentation_b200.inference import InstanceHeat
    dev = torch.device("cuda", torch.cuda.current_device())
    sd = synthetic.make_state_dict(seed=0)
    engine = InstanceHeat(model=None, precision=args.precision, device=dev)
    engine.model.load_state_dict(sd, strict=True)
    engine.packed_k = MAX_DETS
    del sd
    torch.manual_seed(0)
    # the input crosses PCIe as the camera delivers it: uint8 HWC (cv2 BGR), 3 B / pixel; normalisation (test.py:92) runs on the device
    x_host = torch.randint(0, 256, (BS, HW_IN, HW_IN, 3), dtype=torch.uint8).pin_memory()
    x_dev = x_host.to(dev)
    base, host = planted_batch()
    forced = [tuple(torch.from_numpy(a).to(dev) for a in h) for h in host] if not args.free_running else None
    det_host = torch.empty(BS, MAX_DETS + 1, 5, dtype=torch.float64).pin_memory()
    mask_host = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    state = {}
    # the single collective of the path: ONE all-gather (NCCL) of the fixed-size per-image detection records
    # [B_local, MAX_DETS + 1, 5] f64 (row 0 = count), issued on a side stream right after the decode so that it overlaps
    # forward_seg and the next batch.  No slice copy, no second collective for the counts.
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    gathered = torch.empty(world * BS, MAX_DETS + 1, 5, dtype=torch.float64, device=dev) if world > 1 else None
    decoded, comm_done = torch.cuda.Event(), torch.cuda.Event()

    def on_decoded(res):
        if world > 1:
            cur = torch.cuda.current_stream()
            decoded.record(cur)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(decoded)
                dist.all_gather_into_tensor(gathered, res.packed)
                comm_done.record(comm_stream)

    # Two batches in flight (InstanceHeat.submit / collect): forward_dec + decode of batch i+1 are enqueued BEFORE the host waits for
    # the boxes of batch i and plans its forward_seg, so the device never idles on the host round trip.  Every step submits one batch
    # and collects one: K timed steps contain K forward_dec + decode and K forward_seg.  --serial times detect_batch instead.
    def step_device():
        if world > 1:
            torch.cuda.current_stream().wait_event(comm_done)      # the previous gather has read the record buffer (long done)
        if args.serial:
            dets, seg = engine.detect_batch(x_dev, head_override=forced, packed=True, on_decoded=on_decoded)
        else:
            engine.submit(x_dev, head_override=forced, on_decoded=on_decoded)
            if engine._n_submitted - engine._n_collected < 2:
                return None
            dets, seg = engine.collect(packed=True)
        state["dets"] = dets
        return dets

    def drain():
        while engine._n_submitted > engine._n_collected:
            state["dets"] = engine.collect(packed=True)[0]

    # e2e: every step's input crosses PCIe inside the timed region.  The copy of step i+1 runs on a side stream while step i
    # computes (two staging buffers); detections + mask patches of every step are copied back before the step ends.
    copy_stream = torch.cuda.Stream(device=dev)
    x_stages = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_h2d(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])                               # the step that read this buffer has finished
            x_stages[b].copy_(x_host, non_blocking=True)                      # H2D of step i's input from pinned memory
            copied[b].record(copy_stream)

    landed = [torch.cuda.Event(), torch.cuda.Event()]
    produced = torch.cuda.Event()
    d2h_stream = torch.cuda.Stream(device=dev)

    def read_back(i):
        """D2H of the batch's result (detection records + mask patches) on its own stream: the 16 MB of masks would otherwise hold
        the compute stream for 0.6 ms per step.  landed[i & 1] = the copies of batch i are on the host."""
        cur = torch.cuda.current_stream()
        packed = engine.last_result.packed
        m = engine.last_seg_model.last_masks
        nf = min(m.numel(), mask_host.numel())
        produced.record(cur)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(produced)
            det_host.copy_(packed, non_blocking=True)
            mask_host[:nf].copy_(m[:nf], non_blocking=True)
            m.record_stream(d2h_stream)                                       # the caching allocator must not recycle it before the copy ran
            landed[i & 1].record(d2h_stream)
        state["mask_floats"] = nf

    def run_e2e(steps):
        cur = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(cur)
        issue_h2d(0)
        done = 0
        for i in range(steps):
            b = i & 1
            cur.wait_event(copied[b])
            if i + 1 < steps:
                issue_h2d(i + 1)
            if world > 1:
                cur.wait_event(comm_done)
            if args.serial:
                engine.detect_batch(x_stages[b], head_override=forced, packed=True, on_decoded=on_decoded)
                consumed[b].record(cur)
                read_back(i)
                landed[i & 1].synchronize()
                continue
            cur.wait_event(landed[i & 1])          # this slot's record buffer (batch i - 2) has been read back (long done)
            engine.submit(x_stages[b], head_override=forced, on_decoded=on_decoded)
            consumed[b].record(cur)
            if engine._n_submitted - engine._n_collected == 2:
                engine.collect(packed=True)
                read_back(done)
                if done > 0:
                    landed[(done - 1) & 1].synchronize()                      # the previous batch's results are on the host (bounded lag)
                done += 1
        while engine._n_submitted > engine._n_collected:                      # the last batch in flight belongs to the timed region
            engine.collect(packed=True)
            read_back(done)
            done += 1
        cur.wait_event(landed[0]); cur.wait_event(landed[1])                  # every read-back lands inside the timed region
        cur.synchronize()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if world > 1:
            torch.cuda.current_stream().wait_event(comm_done)     # the last gather belongs to the timed region
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    for _ in range(warm + (0 if args.serial else 1)):
        step_device()
    n_det = sum(0 if d is None else len(d) for d in state["dets"])
    truncated = bool(int(engine.last_result.status.item()) & 4)
    launches_per_step = engine.last_launches
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    ms = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else None
    drain()
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)
    if world > 1:     # the gathered records must equal every rank's own detections, ordered by global image index
        torch.cuda.synchronize()
        mine = engine.last_result.packed
        assert torch.equal(gathered[rank * BS:(rank + 1) * BS], mine), "all-gather returned a different detection record"

    # per-stage device time, live, CUDA events on the launching stream (separate pass: events serialise nothing but
    # add host work, so they are kept out of the headline timing)
    _cabi.timing_enable(True)
    for _ in range(args.steps):
        step_device()
    drain()
    st_ms, st_cnt = _cabi.timing_collect()
    _cabi.timing_enable(False)
    # device idle time between the last decode kernel and the first forward_seg launch (D2H of the boxes, host unpack, atlas planning)
    engine.gap_events = []
    for _ in range(args.steps):
        state["dets"] = engine.detect_batch(x_dev, head_override=forced, packed=True)[0]     # serial path: what submit / collect hide
    torch.cuda.synchronize()
    gaps = [a.elapsed_time(b) for a, b in engine.gap_events]
    engine.gap_events = None
    info = np.zeros(8, np.float64)
    _cabi.check(_cabi.lib().kg_net_plan_info(engine.model._handle, info.ctypes.data, 8))
    stages = {STAGE_NAMES[i]: {"ms_per_step": round(float(st_ms[i]) / args.steps, 4), "launches_per_step": int(st_cnt[i]) // args.steps}
              for i in STAGE_NAMES if st_cnt[i] > 0}
    tc_ms = float(st_ms[9:13].sum()) / args.steps
    tc_launches = int(st_cnt[9:13].sum()) // args.steps
    pk, pk_kind = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    achieved = info[0] / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    per_class = {}
    for j, nme in enumerate(("tc_backbone", "tc_decoder", "tc_heads_l1", "tc_heads_l2")):
        t = float(st_ms[9 + j]) / args.steps
        if t > 0:
            per_class[nme] = {"tflops": round(info[4 + j] / (t * 1e-3) / 1e12, 1), "frac": round(info[4 + j] / (t * 1e-3) / 1e12 / peak_tf, 4)}
    sflops, sboxes = seg_flops(state["dets"], HW_IN, HW_IN)
    seg_ms = float(st_ms[16]) / args.steps
    if seg_ms > 0:
        per_class["forward_seg"] = {"tflops": round(sflops / (seg_ms * 1e-3) / 1e12, 1), "frac": round(sflops / (seg_ms * 1e-3) / 1e12 / peak_tf, 4),
                                    "algorithmic_flops_per_step": sflops, "boxes": sboxes}
    whole = (info[0] + info[1] + sflops) / (ms / args.steps * 1e-3) / 1e12
    traffic, traffic_src = profiled_traffic()
    roofline = {"kernel": "tc_conv_kernel + tc_shift_kernel (tcgen05 implicit-GEMM / row-GEMM shift-add convs: all tensor-core launches of forward_dec)",
                "bound": "tensor", "achieved": round(achieved, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4),
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": pk_kind + " (cuBLAS bf16 sustained; kernel timed inside a long step)",
                "algorithmic_flops_per_step": info[0], "launches_per_step": tc_launches, "avg_launch_ms": round(tc_ms / max(1, tc_launches), 4),
                "cuda_core_conv_flops_per_step": info[1], "whole_step_tflops": round(whole, 1), "whole_step_frac": round(whole / peak_tf, 4),
                "per_class": per_class, "stages": stages,
                "serial_host_gap_ms_per_step": round(float(np.mean(gaps)), 4) if gaps else None}
    gbs = world * BS
    out = {
        "metric": METRIC, "value": round(gbs * args.steps / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "fp16 operands (split hi+lo x3 in backbone/decoder, x2 in three 64-channel decoder convs, single pass in heads), fp32 accumulate; fp64 decode"
                 if args.precision == "fast" else args.precision,
        "data": f"synthetic (seeded Kaiming weights with calibrated heads; uniform uint8 images; planted {CELLS}-cell scenes)",
        "config": {"workload": f"bs{BS}/GPU {HW_IN}x{HW_IN}: uint8 HWC -> normalise -> forward_dec (ResNet-50 trunk + decoder + 12 heads) -> vote/blur/peak/group/boxes/NMS -> forward_seg"
                               + (f"; decode teacher-forced with planted {CELLS}-cell head maps (SURVEY.md 8d-ii)" if forced is not None else "; free-running decode (the network's own head maps)"),
                   "config": args.config, "global_batch": gbs, "precision": args.precision, "detections_per_step": n_det,
                   "detection_record_truncated": truncated,
                   "batches_in_flight": 1 if args.serial else 2,
                   "env_switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("KG_")},
                   "l2_note": "activations of one step (>20 GB) exceed L2; no flush needed"},
        "e2e": {"value": round(gbs * args.steps / (ms_e2e * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": x_host.numel(),
                "d2h_bytes_per_step": det_host.numel() * 8 + int(state.get("mask_floats", 0)) * 4},
        "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_pipeline(base, n_images=1, warm=0)
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        del engine, forced
        torch.cuda.empty_cache()
        out["gpu_library_baseline"] = gpu_library_baseline(dev)
        fd_ms = sum(v["ms_per_step"] for k, v in stages.items() if k in ("conv_cuda_core", "tc_backbone", "tc_decoder", "tc_heads_l1",
                                                                         "tc_heads_l2", "bilinear", "maxpool"))
        out["gpu_library_baseline"]["own_forward_dec_ms_per_step"] = round(fd_ms, 3)
    return out


def cpu_pipeline(base, n_images, warm=0):
    """The reference path on host cores via the oracle port: forward_dec (torch fp32, all cores) -> 4x decode -> refine ->
    gather -> NMS -> forward_seg, one image at a time like test.py:88-125."""
    import torch
    from oracle import kg_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(seed=0)
    torch.manual_seed(0)
    xs = torch.rand(max(1, n_images), 3, HW_IN, HW_IN) - 0.5
    def one(i):
        out = O.forward_dec(sd, xs[i % len(xs)][None])
        det, _, _ = O.decode_image(base[i % len(base)])
        O.forward_seg(sd, out[4], [det if det is not None else []])
    for i in range(warm):
        one(i)
    t0 = time.time()
    for i in range(n_images):
        one(i)
    dt = time.time() - t0
    return {"value": round(n_images / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_images} image(s) {HW_IN}x{HW_IN} of the same workload, bs=1 like test.py: torch fp32 forward_dec on {cores} threads + "
                      "NumPy decode (1 thread) + forward_seg (oracle/kg_oracle.py)", "seconds": round(dt, 2)}


def cpu_baseline_decode(base, budget_s=12.0):
    from oracle import kg_oracle as O
    t0 = time.time(); n = 0
    while True:
        O.decode_image(base[n % len(base)])
        n += 1
        if time.time() - t0 > budget_s or n >= 16:
            break
    dt = time.time() - t0
    return {"value": round(n / dt, 4), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} image(s) of the same planted 512x512 workload, NumPy oracle decode (oracle/kg_oracle.py), single thread"}


def reference_modules():
    """The UNMODIFIED reference modules (KGnet, postprocessing, nms) when they can be imported: from baseline/_ref/ (git-ignored copy
    made by baseline/fetch_ref.py; it travels to the GPU box) or /root/reference (this container); None otherwise."""
    import importlib
    if os.environ.get("KG_REFERENCE_PORT"):          # force the oracle port (what a GPU box without the reference runs)
        return None
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if all(os.path.exists(os.path.join(d, f)) for f in ("KGnet.py", "postprocessing.py", "nms.py", "config.py")):
            sys.dont_write_bytecode = True
            sys.path.insert(0, d)
            try:
                for name in ("config", "nms", "postprocessing", "KGnet"):
                    sys.modules.pop(name, None)
                mods = tuple(importlib.import_module(n) for n in ("KGnet", "postprocessing", "nms"))
                if os.path.dirname(os.path.abspath(mods[0].__file__)) == os.path.abspath(d):
                    return mods + (d,)
            except Exception:
                pass
            finally:
                sys.path.remove(d)
    return None


def reference_pipeline(mods, base, n_images, warm=0):
    """test.py:88-125 with the reference's own functions on host cores, one image at a time: forward_dec (torch fp32 CPU) -> 4 x
    get_skeletons_and_masks -> refine_skeleton -> gather_skeleton -> NMS -> forward_seg.  Like the own arm, the decode is fed the
    planted head maps (teacher-forced decode load); weights = the same seeded state dict."""
    import torch
    from kg_instance_segmentation_b200 import synthetic
    KG, PP, NMS, where = mods
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = KG.resnet50(pretrained=False)
    model.load_state_dict(synthetic.make_state_dict(seed=0), strict=True)
    model.eval()
    torch.manual_seed(0)
    xs = torch.rand(max(1, n_images), 3, HW_IN, HW_IN) - 0.5

    def one(i):
        with torch.no_grad():
            out = model.forward_dec(xs[i % len(xs)][None])
        heads = base[i % len(base)]
        sk = [PP.refine_skeleton(PP.get_skeletons_and_masks(*[torch.from_numpy(a[None]) for a in heads[s]])) for s in range(4)]
        boxes = NMS.non_maximum_suppression_numpy(PP.gather_skeleton(*sk), nms_thresh=0.5)
        if boxes is not None:
            with torch.no_grad():
                model.forward_seg(out[4], [boxes])

    for i in range(warm):
        one(i)
    t0 = time.time()
    for i in range(n_images):
        one(i)
    dt = time.time() - t0
    return {"value": round(n_images / dt, 4), "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{n_images} image(s) {HW_IN}x{HW_IN} of the same workload, bs=1 like test.py, through the UNMODIFIED reference modules "
                      f"({where}): forward_dec on {cores} threads + get_skeletons_and_masks x4 + refine + gather + NMS + forward_seg",
            "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores -- the unmodified reference modules when
    they are importable (baseline/_ref/ or /root/reference), else the oracle port (bit-exact restatement,
    tests/test_oracle_vs_reference.py)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    base, _ = planted_batch(n_distinct=2)
    steps = min(args.steps, 6)           # bounded: ~5 s of CPU work per step
    mods = reference_modules()
    if mods is not None:
        r = reference_pipeline(mods, base, n_images=steps, warm=min(args.warmup, 1))
    else:
        r = cpu_pipeline(base, n_images=steps, warm=min(args.warmup, 1))
    v = r["value"]
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": round(1e3 / v, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 (torch CPU) + fp64 decode", "data": "synthetic",
            "config": {"workload": "same path as the own arm, one 512x512 image per step (the reference is batch-size-1 by construction, "
                                   "postprocessing.py:138-140); " + ("unmodified reference modules" if r["kind"] == "reference"
                                                                     else "oracle port of the reference (reference modules not importable)"),
                       "global_batch": 1},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact", "reference"])
    ap.add_argument("--free-running", action="store_true", help="decode the network's own head outputs instead of planted maps")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration (cfg2 = the headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: bs per GPU fixed; strong: the global batch of the config divided over the ranks")
    ap.add_argument("--serial", action="store_true", help="one batch in flight (detect_batch) instead of submit / collect with two")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch/cuDNN forward_dec line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    set_config(args.config, world, args.scaling)
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        # NCCL's version banner / debug log must not share stdout with the JSON line: NCCL honours NCCL_DEBUG_FILE only above
        # the VERSION level, so VERSION (or unset) is raised to WARN and the log goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        torch.cuda.set_device(0)
    from kg_instance_segmentation_b200 import _cabi
    _cabi.lib()   # fail loudly when the CUDA library is missing
    workload = "pipeline" if args.workload == "auto" else args.workload
    if workload == "pipeline":
        out = run_pipeline(args, rank, world, dist)
    else:
        out = run_decode(args, rank, world, dist)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
