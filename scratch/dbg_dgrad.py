import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl4nav_b200 import kernels
dev = "cuda"
g = torch.Generator().manual_seed(1)
def run(B, scale, sparse, label):
    H = W = 20; Cin = 32; Cout = 64; KH = KW = 4; stride = 2; pad = 0
    w = torch.randn(Cout, Cin, KH, KW, generator=g) / 16
    dy = scale * torch.randn(B, 9, 9, Cout, generator=g)
    if sparse: dy = dy * (torch.rand(B, 9, 9, Cout, generator=g) < 0.5)
    mask = torch.randn(B, H, W, Cin, generator=g)
    ref = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), stride=stride).permute(0, 2, 3, 1)
    ref = ref * torch.where(mask > 0, 1.0, 0.01).double()
    for mode in ("tc", "tc2"):
        for rep in range(2):
            out = kernels.conv_nhwc(1, (B, H, W), w.to(dev), dy=dy.to(dev), stride=stride, pad=pad, act=4, mask=mask.to(dev), mode=mode)
            e = (out.cpu().double() - ref).abs()
            bad = (e > 1e-4 * ref.abs().max()).nonzero()
            print(label, "B", B, mode, "rep", rep, "err %.2e" % float(e.max() / ref.abs().max()), "bad", len(bad), bad[:4].tolist())
for B in (8, 5, 16, 300):
    run(B, 1.0, False, "randn")
    run(B, 1e-6, True, "tiny sparse")
