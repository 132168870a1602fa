"""Wall-clock phases of one pipeline step (with syncs) to expose host-side overheads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kg_instance_segmentation_b200 import synthetic
from kg_instance_segmentation_b200.inference import InstanceHeat
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

eng = InstanceHeat(precision="fast")
eng.model.load_state_dict(synthetic.make_state_dict(seed=0), strict=True)
torch.manual_seed(0)
x = (torch.rand(32, 3, 512, 512) - 0.5).cuda()
base, host = bench.planted_batch()
forced = [tuple(torch.from_numpy(a).cuda() for a in h) for h in host]
m = eng.model
m.export_feats = False
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = m.forward_dec(x)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    dec = eng._decoder(32, [tuple(h[0].shape[2:]) for h in forced], 0.5, 4096, 4096)
    res = dec(forced)
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    dets = res.detections()
    t5 = time.perf_counter()
    seg = m.forward_seg(out[4], [d if d is not None else [] for d in dets])
    t6 = time.perf_counter(); torch.cuda.synchronize(); t7 = time.perf_counter()
    print(f"it{it}: fwd_dec host {1e3*(t1-t0):.2f} gpu-wait {1e3*(t2-t1):.2f} | decode host {1e3*(t3-t2):.2f} wait {1e3*(t4-t3):.2f} | "
          f"detections() {1e3*(t5-t4):.2f} | seg host {1e3*(t6-t5):.2f} wait {1e3*(t7-t6):.2f} | total {1e3*(t7-t0):.2f}", flush=True)
