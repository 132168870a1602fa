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

METRIC = "images/sec at 512x512 bs32 (KGnet inference hot path)"
UNIT = "images/s"
BS, HW_IN, CELLS = 32, 512, 40
MAX_PEAKS, MAX_BOXES, MAX_DETS = 4096, 4096, 512
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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def planted_batch(n_distinct=4):
    """Teacher-forced decode load: planted 40-cell scenes, bs 32 built from n_distinct scenes."""
    from oracle import kg_oracle as O   # input GENERATOR only (synthetic data), never on the measured path
    base = [O.planted_scene(100 + i, HW_IN, HW_IN, CELLS)[0] for i in range(n_distinct)]
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
    dec = postprocessing.Decoder(BS, shapes, max_peaks=MAX_PEAKS, max_boxes=MAX_BOXES)
    stage_in = [tuple(torch.empty_like(t) for t in h) for h in dev_heads]
    out_host = torch.empty(BS, MAX_DETS, 5, dtype=torch.float64).pin_memory()
    cnt_host = torch.empty(BS, dtype=torch.int32).pin_memory()
    gathered = torch.empty(world * BS, MAX_DETS, 5, dtype=torch.float64, device=dev) if world > 1 else None
    gathered_cnt = torch.empty(world * BS, dtype=torch.int32, device=dev) if world > 1 else None

    def step_device():
        r = dec(dev_heads)
        if world > 1:   # the single collective of the path: all-gather of the padded detection list
            dist.all_gather_into_tensor(gathered, r.dets[:, :MAX_DETS].contiguous())
            dist.all_gather_into_tensor(gathered_cnt, r.det_count)
        return r

    def step_e2e():
        for hs, ds in zip(host_pinned, stage_in):
            for h, d in zip(hs, ds):
                d.copy_(h, non_blocking=True)
        r = dec(stage_in)
        if world > 1:
            dist.all_gather_into_tensor(gathered, r.dets[:, :MAX_DETS].contiguous())
            dist.all_gather_into_tensor(gathered_cnt, r.det_count)
        out_host.copy_(r.dets[:, :MAX_DETS], non_blocking=True)
        cnt_host.copy_(r.det_count, non_blocking=True)
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
    d2h = out_host.numel() * 8 + cnt_host.numel() * 4

    # per-kernel time of the dominant kernel, live, with CUDA events on the launching stream
    _cabi.timing_enable(True)
    for _ in range(args.steps):
        step_device()
    st_ms, st_cnt = _cabi.timing_collect()
    _cabi.timing_enable(False)
    px = BS * sum(h * w for h, w in shapes)
    names = {0: "vote_kernel", 1: "blur_peak_kernel", 2: "group_kernel", 3: "nms_kernel"}
    stage = {names[i]: {"ms_per_step": float(st_ms[i]) / args.steps, "launches_per_step": int(st_cnt[i]) // args.steps} for i in names}
    dom = max(names, key=lambda i: st_ms[i])
    alg_bytes = px * (VOTE_BYTES_PX if dom == 0 else BLUR_BYTES_PX if dom == 1 else 0)
    pk, pk_kind = peaks()
    dom_ms = float(st_ms[dom]) / args.steps
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"kernel": names[dom], "bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": None, "peak_kind": pk_kind + " (burst copy)",
                "algorithmic_bytes_per_launch": alg_bytes // max(1, stage[names[dom]]["launches_per_step"]),
                "stages": stage}
    out = {
        "metric": METRIC, "value": round(world * BS * args.steps / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (planted 40-cell scenes, teacher-forced head maps)",
        "config": {"workload": "decode-only: bs32/GPU 512x512 head maps (4 scales) -> vote+blur+peak+group+boxes+NMS; inputs 2.45 GB > L2, no flush needed",
                   "global_batch": world * BS, "detections_per_step": n_det},
        "e2e": {"value": round(world * BS * args.steps / (ms_e2e * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_decode(base)
    return out


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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference is pure Python and cannot
    travel to the GPU box, so this times the oracle port (bit-exact restatement, tests/test_oracle_vs_reference.py)."""
    import torch
    from oracle import kg_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    base, _ = planted_batch()
    per_step = 1
    for _ in range(min(args.warmup, 1)):
        O.decode_image(base[0])
    t0 = time.time()
    for k in range(args.steps):
        for j in range(per_step):
            O.decode_image(base[(k + j) % len(base)])
    dt = time.time() - t0
    v = round(args.steps * per_step / dt, 4)
    sample = f"{per_step} image per step (bounded sample of the bs32 512x512 planted workload), decode-only, NumPy oracle port"
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "decode-only (same as the own arm)", "global_batch": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    else:
        torch.cuda.set_device(0)
    from kg_instance_segmentation_b200 import _cabi
    _cabi.lib()   # fail loudly when the CUDA library is missing
    workload = args.workload
    if workload == "auto":
        try:
            from kg_instance_segmentation_b200 import KGnet  # noqa: F401
            workload = "pipeline"
        except ImportError:
            workload = "decode"
    if workload == "pipeline":
        from kg_instance_segmentation_b200 import bench_pipeline
        out = bench_pipeline.run(args, rank, world, dist)
    else:
        out = run_decode(args, rank, world, dist)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
