"""Bring-up diagnostics for the tcgen05 conv kernel: error maps against torch fp32 on a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from kg_instance_segmentation_b200 import _cabi

L = _cabi.lib()
print("tc available:", L.kg_tc_available(), L.kg_tc_status().decode())


def run(N, Cin, H, W, Cout, k, mode, relu=False, seed=0, ident=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    if ident:
        w.zero_()
        for c in range(min(Cin, Cout)):
            w[c, c, k // 2, k // 2] = 1.0
    b = torch.zeros(Cout) if ident else torch.randn(Cout, generator=g) * 0.1
    y = F.conv2d(x, w, b, padding=k // 2)
    if relu:
        y = F.relu(y)
    xd = x.cuda(); out = torch.zeros_like(y).cuda()
    rc = L.kg_conv2d_nchw(xd.data_ptr(), N, Cin, H, W, w.data_ptr(), b.data_ptr(), Cout, k, k, 1, k // 2, int(relu), None, mode,
                          out.data_ptr(), None)
    if rc != 0:
        print(f"  N{N} C{Cin}->{Cout} {H}x{W} k{k} mode{mode}: ERROR {rc} {L.kg_last_error().decode()}")
        return
    d = (out.cpu() - y).abs()
    print(f"  N{N} C{Cin}->{Cout} {H}x{W} k{k} mode{mode} ident={ident}: max err {d.max():.3e} (scale {y.abs().max():.2f}) "
          f"bad frac {(d > 1e-2 * y.abs().max()).float().mean():.4f}")
    if d.max() > 1e-2 * y.abs().max():
        per_c = d.amax(dim=(0, 2, 3)); per_y = d.amax(dim=(0, 1, 3)); per_x = d.amax(dim=(0, 1, 2))
        print("    per-channel(16) :", [f"{v:.2f}" for v in per_c[:16].tolist()])
        print("    per-row(16)     :", [f"{v:.2f}" for v in per_y[:16].tolist()])
        print("    per-col(16)     :", [f"{v:.2f}" for v in per_x[:16].tolist()])
        print("    out[0,0,0,:8]   :", out[0, 0, 0, :8].cpu().tolist())
        print("    ref[0,0,0,:8]   :", y[0, 0, 0, :8].tolist())


for mode in (1, 3):
    run(1, 64, 16, 8, 64, 1, mode)
    run(1, 64, 16, 16, 64, 3, mode)
    run(1, 64, 8, 128, 64, 3, mode, ident=True)     # strip mode (W >= 128)
    run(1, 64, 8, 128, 64, 3, mode)
    run(2, 64, 4, 256, 64, 3, mode, relu=True)
    run(1, 64, 6, 128, 192, 7, mode)
    run(1, 128, 5, 384, 128, 3, mode)
    run(1, 256, 16, 16, 512, 3, mode)
torch.cuda.synchronize()
print("done")
