"""Max |kp heatmap| error of the precision modes against the 3-pass 'exact' mode at benchmark resolution."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kg_instance_segmentation_b200 import synthetic, KGnet

sd = synthetic.make_state_dict(seed=0)
m = KGnet.resnet50(pretrained=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
m.export_feats = False
torch.manual_seed(0)
x = (torch.rand(2, 3, 512, 512) - 0.5).cuda()
m.precision = "exact"
ref = [[t.clone() for t in o] for o in m.forward_dec(x)[:4]]
for prec in ("reference", "fast"):
    m.precision = prec
    out = m.forward_dec(x)[:4]
    errs = [float((out[s][0] - ref[s][0]).abs().max()) for s in range(4)]
    offs = [float((out[s][k] - ref[s][k]).abs().max()) for s in range(4) for k in (1, 2)]
    print(prec, "kp err per scale", ["%.2e" % e for e in errs], "max offset err %.2e" % max(offs), flush=True)
