"""Host-side cost of one pipeline step (diagnostics): python tools/host_profile.py [bs] [hw]
cProfile over 20 submit / collect steps at a small batch, where the step is host-bound (strong scaling at 8 GPUs: bs4 per rank)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kg_instance_segmentation_b200 import synthetic
from kg_instance_segmentation_b200.inference import InstanceHeat

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 512
eng = InstanceHeat(model=None, precision="fast", device="cuda:0")
eng.model.load_state_dict(synthetic.make_state_dict(seed=0), strict=True)
eng.packed_k = 1024
x = torch.randint(0, 256, (bs, hw, hw, 3), dtype=torch.uint8, device="cuda")
scenes = [synthetic.planted_scene(100 + i, hw, hw, 40)[0] for i in range(min(bs, 4))]
forced = [tuple(torch.from_numpy(np.stack([scenes[i % len(scenes)][s][k] for i in range(bs)])).cuda() for k in range(3)) for s in range(4)]


def step():
    eng.submit(x, head_override=forced)
    if eng._n_submitted - eng._n_collected == 2:
        eng.collect(packed=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
torch.cuda.synchronize()
print(f"bs{bs}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms per step (wall)")
# host time alone: let the device drain before every step so that nothing blocks on it
pr = cProfile.Profile()
host = 0.0
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr.enable()
    step()
    pr.disable()
    host += time.perf_counter() - t0
print(f"host time per step with an idle device: {host / 20 * 1e3:.2f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
