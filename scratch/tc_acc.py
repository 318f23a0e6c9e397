import sys; sys.path.insert(0,'/root/repo')
import torch
from ddrl4nav_b200 import kernels
dev='cuda'
def tf32_exact(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
g=torch.Generator().manual_seed(0)
for K in [256,1024,4096,16384]:
    A=torch.randn(256,K,generator=g); B=torch.randn(128,K,generator=g)
    for name,(a,b) in {"general":(A,B),"tf32exact":(tf32_exact(A),tf32_exact(B)),"positive":(A.abs(),B.abs()),"pos_exact":(tf32_exact(A.abs()),tf32_exact(B.abs()))}.items():
        ref=a.double()@b.double().T
        tc=kernels.gemm(0,a.to(dev),b.to(dev),mode='tc').cpu().double()
        si=kernels.gemm(0,a.to(dev),b.to(dev),mode='simt').cpu().double()
        f32=(a.to(dev)@b.to(dev).T).cpu().double()
        mx=ref.abs().max()
        rel=lambda x: float(((x-ref)/ref.abs().clamp_min(1e-30)).abs().median())
        print(K,name,"tc max/mx %.2e med-rel %.2e signed-mean-rel %.2e | simt %.2e %.2e | cublas-fp32 %.2e"%(float((tc-ref).abs().max()/mx),rel(tc),float(((tc-ref)/ref.abs().clamp_min(1e-30)).mean()),float((si-ref).abs().max()/mx),rel(si),float((f32-ref).abs().max()/mx)))
