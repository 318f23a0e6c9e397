"""Runs a few launches of one implicit conv (for ncu captures). usage: one_conv.py mode op B H W Cin Cout KH KW stride pad"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl4nav_b200 import kernels
mode, op = sys.argv[1], int(sys.argv[2])
B, H, W, Cin, Cout, KH, KW, stride, pad = [int(x) for x in sys.argv[3:12]]
dev = "cuda"
x = torch.randn(B, H, W, Cin, device=dev)
w = torch.randn(Cout, Cin, KH, KW, device=dev)
Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
dy = torch.randn(B, Ho, Wo, Cout, device=dev)
for _ in range(3):
    out = kernels.conv_nhwc(op, x if op != 1 else (B, H, W), w, dy=dy if op else None, stride=stride, pad=pad, mode=mode)
torch.cuda.synchronize()
print("done", out.shape)
