#!/usr/bin/env python
"""Role-level accounting of the tc3 forward kernel (needs the -DTC3_TIMING build):
   make -C ddrl4nav_b200/csrc timing3
   DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing.so python scratch/tc3_roles.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl4nav_b200 import _lib, kernels

lib = _lib.load()
lib.ddrl_tc3_timing_read.restype = C.c_int
lib.ddrl_tc3_timing_read.argtypes = [C.c_void_p, C.c_int]
dev = "cuda"
WNAMES = {5: ("w-producer", ["empty"]), 6: ("w-mma-chunk", ["mfree", "aready", "issue+commit"]),
          7: ("w-splitter", ["full", "afree", "gather+split+st+wait", "of which st+wait::st+arrive"]), 8: ("w-epilogue", ["mfull", "cfull", "atomics"])}
NAMES = {0: ("producer", ["empty"]), 1: ("mma-chunk", ["mfree", "aready", "issue+commit"]), 2: ("mma-corr", ["cfree", "aready", "issue+commit"]),
         3: ("splitter", ["full", "afree", "lds+split+st (incl. afree)"]), 4: ("epilogue", ["mfull", "cfull", "stores"])}


def report(tag, fn, reps=3, names=None):
    names = names or NAMES
    fn(); torch.cuda.synchronize()
    lib.ddrl_tc3_timing_read(None, 1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 96)()
    lib.ddrl_tc3_timing_read(buf, 1)
    print("== %s" % tag)
    for role, (nm, ws) in names.items():
        life = buf[role * 8 + 7] or 1
        print("   %-10s lifetime %8.0f kcyc/CTA-launch | " % (nm, life / 1e3 / reps / 148) +
              "  ".join("%s %4.1f%%" % (w, 100.0 * buf[role * 8 + k] / life) for k, w in enumerate(ws)))


B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
def conv(op, H, W, Cin, Cout, K, s, pad=0):
    x = torch.randn(B, H, W, Cin, device=dev); w = torch.randn(Cout, Cin, K, K, device=dev)
    Ho, Wo = (H + 2 * pad - K) // s + 1, (W + 2 * pad - K) // s + 1
    dy = torch.randn(B, Ho, Wo, Cout, device=dev)
    return lambda: kernels.conv_nhwc(op, x if op != 1 else (B, H, W), w, dy=dy if op else None, stride=s, pad=pad, mode="tc3")
def gemm(form, M, N, K):
    if form == 0: A, Bm = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    elif form == 1: A, Bm = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev)
    else: A, Bm = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev)
    return lambda: kernels.gemm(form, A, Bm, mode="tc3")

report("conv2 fwd 20x20x32 -> 9x9x64 k4 s2", conv(0, 20, 20, 32, 64, 4, 2))
report("conv3 fwd 9x9x64 -> 7x7x64 k3 s1", conv(0, 9, 9, 64, 64, 3, 1))
report("conv3 dgrad", conv(1, 9, 9, 64, 64, 3, 1))
report("conv2 dgrad (fused parity)", conv(1, 20, 20, 32, 64, 4, 2))
report("conv1 gemm fwd M=B*400 N=64 K=256", gemm(0, B * 400, 64, 256))
report("fc fwd M=B N=512 K=3136", gemm(0, B, 512, 3136))

report("conv2 wgrad 20x20x32 -> 9x9x64 k4 s2", conv(2, 20, 20, 32, 64, 4, 2), names=WNAMES)
report("conv3 wgrad 9x9x64 -> 7x7x64 k3 s1", conv(2, 9, 9, 64, 64, 3, 1), names=WNAMES)
report("conv1 (s2d) wgrad 21x21x64 -> 20x20x64 k2 s1", conv(2, 21, 21, 64, 64, 2, 1), names=WNAMES)
report("fc wgrad M=3136 N=512 K=B", gemm(2, 3136, 512, B), names=WNAMES)
