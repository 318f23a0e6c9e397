import sys; sys.path.insert(0,'/root/repo')
import torch
from ddrl4nav_b200 import kernels
dev='cuda'
B=8192
shapes=[ # name, form, M,N,K
 ("conv1 fwd",0,B*400,32,256),("conv2 fwd",0,B*81,64,512),("conv3 fwd",0,B*49,64,576),("linear fwd",0,B,512,3136),
 ("linear dgrad",1,B,3136,512),("conv3 dgrad",1,B*49,576,64),("conv2 dgrad",1,B*81,512,64),
 ("linear wgrad",2,3136,512,B),("conv3 wgrad",2,576,64,B*49),("conv2 wgrad",2,512,64,B*81),("conv1 wgrad",2,256,32,B*400),
 ("big square",0,8192,8192,8192),("nav conv2 fwd",0,1024*400,128,1600),
]
def run(mode,form,M,N,K):
    if form==0: A=torch.randn(M,K,device=dev); Bm=torch.randn(N,K,device=dev)
    elif form==1: A=torch.randn(M,K,device=dev); Bm=torch.randn(K,N,device=dev)
    else: A=torch.randn(K,M,device=dev); Bm=torch.randn(K,N,device=dev)
    out=torch.empty(M,N,device=dev)
    for _ in range(2): kernels.gemm(form,A,Bm,mode=mode,out=out)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    n=5
    e0.record()
    for _ in range(n): kernels.gemm(form,A,Bm,mode=mode,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n
    fl=2.0*M*N*K; by=4.0*(M*K+N*K+M*N)
    return ms, fl/ms/1e9, by/ms/1e6
for name,form,M,N,K in shapes:
    r=[]
    for mode in ["tc","simt"]:
        if mode=="simt" and M*N*K>3e11: r.append((0,0,0)); continue
        r.append(run(mode,form,M,N,K))
    print("%-14s M=%8d N=%5d K=%8d | tc %8.3f ms %7.1f TF/s %7.0f GB/s | simt %8.3f ms %6.1f TF/s"%(name,M,N,K,r[0][0],r[0][1],r[0][2],r[1][0],r[1][1]))
