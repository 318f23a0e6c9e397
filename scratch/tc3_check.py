"""tc3 vs tc2 on the Pong learner's GEMM shapes: accuracy against fp64 and per-launch kernel time (ddrl_prof events)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DDRL_PROF_SHAPES"] = "1"
import torch
from ddrl4nav_b200 import kernels, _lib
lib = _lib.load()
dev = "cuda"
shapes = [(663552, 64, 512), (401408, 64, 576), (819200, 128, 256), (3276800, 64, 256), (8192, 512, 3136), (8192, 3136, 512), (100000, 32, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for M, N, K in shapes:
    g = torch.Generator(device=dev).manual_seed(M % 1000 + N + K)
    A = torch.randn(M, K, device=dev, generator=g)
    B = torch.randn(N, K, device=dev, generator=g) * 0.05
    rows = torch.randint(0, M, (256,), device=dev)
    ref = A[rows].double() @ B.double().T
    line = "M=%d N=%d K=%d:" % (M, N, K)
    for mode in ("tc2", "tc3"):
        out = kernels.gemm(0, A, B, mode=mode)
        err = float((out[rows].double() - ref).abs().max() / ref.abs().max())
        _lib.check(lib.ddrl_prof_start(_lib.current_stream()))
        for _ in range(3):
            kernels.gemm(0, A, B, mode=mode)
        buf = C.create_string_buffer(1 << 16)
        _lib.check(lib.ddrl_prof_stop(buf, 1 << 16))
        t = [l.split() for l in buf.value.decode().splitlines() if "gemm_tc" in l]
        us = sum(float(x[1]) for x in t) / max(1, sum(int(x[2]) for x in t)) * 1e3
        line += "  %s %.1f us (%.0f TF/s) err %.2e" % (mode, us, 2.0 * M * N * K / us / 1e6, err)
        del out
    print(line, flush=True)
    del A, B
