"""CPU emulation of operand rounding (fp16 single-pass) per layer group, to decide which groups may run single-pass
under the 1e-3 heat-map contract.  python tools/precision_emulate.py [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from kg_instance_segmentation_b200 import synthetic

def q16(t):
    return t.half().float()

def run(sd, x, groups, blocks=(3, 4, 6)):
    def conv(xx, name, group, bias=True, stride=1, pad=0):
        w = sd[name + ".weight"]; b = sd.get(name + ".bias") if bias else None
        if (group + ":a") in groups:                          # "<group>:a" rounds only the activations (weights stay split-fp16)
            xx = q16(xx)
        if group in groups or (group + ":w") in groups:       # "<group>:w" rounds only the weights (activations stay split-fp16)
            mx = w.abs().max(); sc = 2.0 ** (10 - torch.frexp(mx)[1].item())
            if group in groups:
                xx = q16(xx)
            w = q16(w * sc) / sc
        return F.conv2d(xx, w, b, stride=stride, padding=pad)
    def bn(xx, p):
        return F.batch_norm(xx, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)
    up = lambda a, ref: F.interpolate(a, ref.shape[2:], mode="bilinear", align_corners=False)
    with torch.no_grad():
        c0 = F.relu(conv(F.relu(conv(x, "c0_conv.0", "stem", pad=1)), "c0_conv.2", "c0b", pad=1))
        c1 = F.relu(bn(conv(x, "conv1", "stem", bias=False, stride=2, pad=3), "bn1"))
        y = F.max_pool2d(c1, 3, 2, 1)
        feats = []
        for li, nb in enumerate(blocks):
            for b in range(nb):
                p = f"layer{li+1}.{b}"; st = 2 if (b == 0 and li > 0) else 1
                o = F.relu(bn(conv(y, p + ".conv1", "backbone", bias=False), p + ".bn1"))
                o = F.relu(bn(conv(o, p + ".conv2", "backbone", bias=False, stride=st, pad=1), p + ".bn2"))
                o = bn(conv(o, p + ".conv3", "backbone", bias=False), p + ".bn3")
                if (p + ".downsample.0.weight") in sd:
                    y = bn(conv(y, p + ".downsample.0", "backbone", bias=False, stride=st), p + ".downsample.1")
                y = F.relu(o + y)
            feats.append(y)
        c2, c3, c4 = feats
        c4u = F.relu(conv(up(c4, c3), "c4_up_conv.0", "dec_up", pad=1))
        c3c = F.relu(conv(torch.cat((c4u, c3), 1), "c3_cat_refine.0", "dec_cat"))
        c3u = F.relu(conv(up(c3c, c2), "c3_up_conv.0", "dec_up", pad=1))
        c2c = F.relu(conv(torch.cat((c3u, c2), 1), "c2_cat_refine.0", "dec_cat"))
        c2u = F.relu(conv(up(c2c, c1), "c2_up_conv.0", "dec_up" if "c2up:w" not in groups else "c2up", pad=1))
        c1c = F.relu(conv(torch.cat((c2u, c1), 1), "c1_cat_refine.0", "dec_cat"))
        c1u = F.relu(conv(up(c1c, c0), "c1_up_conv.0", "dec_up" if "c1up:w" not in groups else "c1up", pad=1))
        c0c = F.relu(conv(torch.cat((c1u, c0), 1), "c0_cat_refine.0", "dec_cat"))
        outs = []
        for s, f in enumerate((c0c, c1c, c2c, c3c)):
            hd = lambda name: conv(F.relu(conv(f, f"{name}_c{s}.0", "head1", pad=3)), f"{name}_c{s}.2", "head2", pad=3)
            outs.append([torch.sigmoid(hd("kp_head")), hd("short_offset_head"), hd("mid_offset_head")])
    return outs

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.set_num_threads(8)
for seed in (0, 1):
    sd = synthetic.make_state_dict(seed=seed)
    torch.manual_seed(seed)
    x = torch.rand(1, 3, size, size) - 0.5
    ref = run(sd, x, set())
    cases = os.environ.get("KG_EMU_CASES")
    for groups in ([c.split("+") for c in cases.split(",")] if cases else []) or (["head1", "head2"], ["head1", "head2", "c0b:w", "c1up:w"], ["head1", "head2", "c0b:w", "c1up:w", "c2up:w"],
                   ["head1", "head2", "c0b:w", "c1up:w", "dec_cat:w"], ["head1", "head2", "dec_up"], ["head1", "head2", "dec_cat"], ["head1", "head2", "dec_up", "dec_cat"],
                   ["head1", "head2", "dec_up", "dec_cat", "c0b"], ["head1", "head2", "backbone"],
                   ["head1", "head2", "dec_up", "dec_cat", "c0b", "backbone"]):
        out = run(sd, x, set(groups))
        kp = [float((out[s][0] - ref[s][0]).abs().max()) for s in range(4)]
        off = max(float((out[s][k] - ref[s][k]).abs().max()) for s in range(4) for k in (1, 2))
        print(seed, "+".join(groups), "kp", ["%.1e" % e for e in kp], "off %.1e" % off, flush=True)
