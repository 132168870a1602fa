"""One forward_dec (+ decode + forward_seg) at the benchmark shape, for ncu: python tools/profile_step.py [steps] [bs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kg_instance_segmentation_b200 import synthetic
from kg_instance_segmentation_b200.inference import InstanceHeat

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
eng = InstanceHeat(precision=os.environ.get("KG_PRECISION", "fast"))
eng.model.load_state_dict(synthetic.make_state_dict(seed=0), strict=True)
torch.manual_seed(0)
x = (torch.rand(bs, 3, 512, 512) - 0.5).cuda()
for _ in range(steps):
    eng.detect_batch(x, with_masks=os.environ.get("KG_MASKS", "0") == "1")
torch.cuda.synchronize()
print("ok")
