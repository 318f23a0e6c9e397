"""Probe of torch's symmetric-memory plumbing on the GPU box (run under torchrun): allocation + rendezvous, peer / multicast
pointers, and timings of NCCL all-reduce against torch's own symmetric-memory all-reduce ops for the learner's gradient size."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 3368000 + 8
t = symm_mem.empty(N, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
if rank == 0:
    print("rank/world", hdl.rank, hdl.world_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hex(hdl.multicast_ptr or 0),
          "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs], "signal_pad_size", hdl.signal_pad_size, "multicast support",
          getattr(hdl, "has_multicast_support", None), flush=True)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


x = torch.randn(N, device=dev)
res = {}
res["nccl_all_reduce_us"] = timed(lambda: dist.all_reduce(x))
t.copy_(x)
for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        res[name + "_us"] = timed(lambda: op(t, "sum", dist.group.WORLD.group_name))
    except Exception as e:  # noqa: BLE001
        res[name + "_us"] = "ERR " + str(e)[:200]
# this library's kernel: multicast path and plain peer path
import ctypes as C
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddrl4nav_b200 import _lib
lib = _lib.load()
f = symm_mem.empty(lib.ddrl_peer_allreduce_flag_bytes() // 4, dtype=torch.int32, device=dev)
f.zero_()
hf = symm_mem.rendezvous(f, dist.group.WORLD.group_name)
torch.cuda.synchronize(); dist.barrier()
bufs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
flags = (C.c_void_p * world)(*[int(p) for p in hf.buffer_ptrs])
seq = [0]
n4 = N // 4 * 4


def own(mc):
    seq[0] += 1
    rc = lib.ddrl_peer_allreduce_f32(bufs, mc, flags, rank, world, 0, n4, seq[0], C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc


if int(hdl.multicast_ptr or 0):
    res["ddrl_peer_allreduce_multicast_us"] = timed(lambda: own(C.c_void_p(int(hdl.multicast_ptr))))
res["ddrl_peer_allreduce_p2p_us"] = timed(lambda: own(None))
if rank == 0:
    print(res, flush=True)
dist.destroy_process_group()
