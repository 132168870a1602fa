"""Per-op device times of forward_dec at the benchmark shape (diagnostics): KG_TIMING_PER_OP=1 python tools/op_times.py [bs]
Prints ms per plan op (index matches the `[op i name]` lines printed with KG_TC_DEBUG=1)."""
import os, sys
os.environ.setdefault("KG_TIMING_PER_OP", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kg_instance_segmentation_b200 import _cabi, synthetic, KGnet

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = KGnet.resnet50(pretrained=False)
m.load_state_dict(synthetic.make_state_dict(seed=0), strict=True)
m = m.cuda().eval()
m.precision = os.environ.get("KG_PRECISION", "fast")
m.export_feats = False
torch.manual_seed(0)
x = (torch.rand(bs, 3, 512, 512) - 0.5).cuda()
for _ in range(2):
    m.forward_dec(x)
torch.cuda.synchronize()
_cabi.timing_enable(True)
steps = 3
for _ in range(steps):
    m.forward_dec(x)
ms, cnt = _cabi.timing_collect(256)
_cabi.timing_enable(False)
tot = 0.0
for i in range(64, 256):
    if cnt[i]:
        print(f"op {i - 64:3d}: {ms[i] / steps:8.3f} ms")
        tot += ms[i] / steps
print(f"total {tot:.3f} ms")
