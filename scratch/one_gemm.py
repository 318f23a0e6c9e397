"""Runs a few launches of one GEMM/conv shape (for ncu captures). usage: one_gemm.py mode M N K [form]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl4nav_b200 import kernels
mode, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
form = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = "cuda"
if form == 0:
    A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
elif form == 1:
    A, B = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev)
else:
    A, B = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev)
for _ in range(3):
    out = kernels.gemm(form, A, B, mode=mode)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = kernels.gemm(form, A, B, mode=mode)
e1.record()
torch.cuda.synchronize()
print("%s form %d M=%d N=%d K=%d: %.4f ms/call (includes tc2's on-the-fly weight split + sync)" % (mode, form, M, N, K, e0.elapsed_time(e1) / 5))
