#!/usr/bin/env python
"""Kernel time of the Pong / nav layer shapes on one tc2 build (DDRL_LIB_PATH selects an experiment build, see T2_EXP in
csrc/tc2.cu).  usage: DDRL_LIB_PATH=... python scratch/tc2_exp.py [B]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DDRL_PROF_SHAPES"] = "1"
import torch
from ddrl4nav_b200 import _lib, kernels

lib = _lib.load()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192


def timed(tag, fn, reps=3):
    fn(); torch.cuda.synchronize()
    lib.ddrl_prof_start(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    for _ in range(reps):
        fn()
    buf = C.create_string_buffer(1 << 16)
    lib.ddrl_prof_stop(buf, len(buf))
    best = None
    for line in buf.value.decode().splitlines():
        name, ms, n, work = line.split()
        if "tc2[" in name:
            best = (name, float(ms) / int(n))
    print("%-44s %-52s %8.1f us" % (tag, best[0] if best else "-", best[1] * 1e3 if best else -1), flush=True)


def conv(op, H, W, Cin, Cout, K, s, pad=0, b=B):
    x = torch.randn(b, H, W, Cin, device=dev); w = torch.randn(Cout, Cin, K, K, device=dev)
    Ho, Wo = (H + 2 * pad - K) // s + 1, (W + 2 * pad - K) // s + 1
    dy = torch.randn(b, Ho, Wo, Cout, device=dev)
    return lambda: kernels.conv_nhwc(op, x if op != 1 else (b, H, W), w, dy=dy if op else None, stride=s, pad=pad, mode="tc2")


def gemm(form, M, N, K):
    if form == 0: A, Bm = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    elif form == 1: A, Bm = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev)
    else: A, Bm = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev)
    return lambda: kernels.gemm(form, A, Bm, mode="tc2")


timed("pong conv1 s2d fwd 21x21x64 k2 -> 32", conv(0, 21, 21, 64, 32, 2, 1))
timed("pong conv1 s2d wgrad", conv(2, 21, 21, 64, 32, 2, 1))
timed("pong conv1 gemm fwd M=B*400 N=32 K=256", gemm(0, B * 400, 32, 256))
timed("pong conv1 gemm wgrad", gemm(2, 256, 32, B * 400))
timed("pong conv2 fwd 20x20x32 -> 9x9x64 k4 s2", conv(0, 20, 20, 32, 64, 4, 2))
timed("pong conv2 dgrad fused", conv(1, 20, 20, 32, 64, 4, 2))
timed("pong conv2 wgrad", conv(2, 20, 20, 32, 64, 4, 2))
timed("pong conv3 fwd 9x9x64 -> 7x7x64 k3", conv(0, 9, 9, 64, 64, 3, 1))
timed("pong conv3 dgrad", conv(1, 9, 9, 64, 64, 3, 1))
timed("pong conv3 wgrad", conv(2, 9, 9, 64, 64, 3, 1))
timed("pong fc fwd M=B N=512 K=3136", gemm(0, B, 512, 3136))
timed("pong fc wgrad M=3136 N=512 K=B", gemm(2, 3136, 512, B))
timed("navlaser conv2 fwd 22x22x64 k5 p1 -> 128", conv(0, 22, 22, 64, 128, 5, 1, 1, b=1024))
timed("navlaser conv2 wgrad", conv(2, 22, 22, 64, 128, 5, 1, 1, b=1024))
timed("navlaser conv2 dgrad", conv(1, 22, 22, 64, 128, 5, 1, 1, b=1024))
