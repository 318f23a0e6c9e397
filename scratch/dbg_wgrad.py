import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl4nav_b200 import kernels
dev = "cuda"
g = torch.Generator().manual_seed(0)
def err(out, ref): return float((out.cpu().double() - ref).abs().max() / ref.abs().max())
for K, M, N in [(3200, 32, 256), (16000, 32, 256), (3200, 64, 512)]:
    for name, A, B in [
        ("randn", torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)),
        ("tiny dy, U01 x", 1e-6 * torch.randn(K, M, generator=g), torch.rand(K, N, generator=g)),
        ("sparse dy", torch.randn(K, M, generator=g) * (torch.rand(K, M, generator=g) < 0.3), torch.rand(K, N, generator=g)),
        ("randn dy, U01 x", torch.randn(K, M, generator=g), torch.rand(K, N, generator=g)),
    ]:
        ref = A.double().T @ B.double()
        for mode in ("tc", "tc2"):
            out = kernels.gemm(2, A.to(dev), B.to(dev), mode=mode)
            print(K, M, N, name, mode, "%.2e" % err(out, ref))
