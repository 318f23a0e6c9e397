"""torchrun check of ddrl_peer_allreduce_f32 at any world size: result against NCCL (allclose: the summation order differs for
W > 2), replicas bit-identical, timings for the learner's gradient sizes."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from ddrl4nav_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
f = symm_mem.empty(lib.ddrl_peer_allreduce_flag_bytes() // 4, dtype=torch.int32, device=dev)
f.zero_()
hf = symm_mem.rendezvous(f, dist.group.WORLD)
flags = (C.c_void_p * world)(*[int(p) for p in hf.buffer_ptrs])
seq = 0
out = {}
for n in (4096, 3368008, 5630000, 12800012):
    g = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hg = symm_mem.rendezvous(g, dist.group.WORLD)
    bufs = (C.c_void_p * world)(*[int(p) for p in hg.buffer_ptrs])
    torch.cuda.synchronize(); dist.barrier()
    n4 = n // 4 * 4
    for name, mc in (("mc", C.c_void_p(int(hg.multicast_ptr)) if int(hg.multicast_ptr or 0) else None), ("p2p", None)):
        if name == "mc" and mc is None:
            continue
        x = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
        ref = x.clone(); dist.all_reduce(ref)
        g.copy_(x)
        seq += 1
        assert lib.ddrl_peer_allreduce_f32(bufs, mc, flags, rank, world, 0, n4, seq, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
        err = float((g[:n4] - ref[:n4]).abs().max() / ref.abs().max())
        gathered = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gathered, g)
        same = all(torch.equal(gathered[0][:n4], t[:n4]) for t in gathered)

        def run():
            global seq
            seq += 1
            lib.ddrl_peer_allreduce_f32(bufs, mc, flags, rank, world, 0, n4, seq, C.c_void_p(torch.cuda.current_stream().cuda_stream))

        def timed(fn, k=30):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k * 1e3
        t_own = timed(run)
        t_nccl = timed(lambda: dist.all_reduce(x))
        out["%d/%s" % (n, name)] = dict(rel_err=err, replicas_identical=same, own_us=round(t_own, 1), nccl_us=round(t_nccl, 1))
    del g, hg
if rank == 0:
    for k, v in out.items():
        print(k, v, flush=True)
dist.destroy_process_group()
