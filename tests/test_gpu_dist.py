"""2-GPU NCCL test of the data-parallel learner (skipped on single-GPU boxes): the sharded learn step with
one all-reduce per iteration must reproduce the single-GPU full-batch result (SURVEY 8e parity target)."""
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import restate as R

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q, kind="pong", env=None, iters=2):
    import os
    os.environ.update(env or {})
    import torch.distributed as tdist
    from ddrl4nav_b200 import dist
    from ddrl4nav_b200.data import Experience
    from ddrl4nav_b200.runner import make_net
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tdist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=3)
    states = R.synth_states(kind, B, seed=4)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=5)
    net = make_net(kind, device=None, TRAINING_ITER_TIME=iters)
    net.load_state_dict(params if rank == 0 else R.init_params(spec, seed=99))   # rank 1 starts different on purpose
    net = net.to(dev)
    net.enable_data_parallel()
    net.broadcast_parameters(0)
    lo, hi = dist.shard_rows(B, rank, world)
    exp = Experience(states=[s[lo:hi].numpy() for s in states], advs=adv[lo:hi].numpy(), actions=a[lo:hi].numpy(),
                     old_logps=old[lo:hi].numpy(), values=ret[lo:hi].numpy()[None])
    exp.to_tensor(device=dev)
    logs = [l for l, _, _ in net.learn(exp)]
    flat = net._flat.detach().cpu()
    used = "peer-mc" if net._peer and net._peer["mc"] else ("peer" if net._peer else "nccl")
    if rank == 0:
        # single-GPU full batch on the same device
        ref = make_net(kind, device=None, TRAINING_ITER_TIME=iters)
        ref.load_state_dict(params)
        ref = ref.to(dev)
        full = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                          values=ret.numpy()[None])
        full.to_tensor(device=dev)
        rlogs = [l for l, _, _ in ref.learn(full)]
        q.put((logs, rlogs, flat.numpy(), ref._flat.detach().cpu().numpy(), used))
    else:
        q.put(("rank1", flat.numpy()))
    tdist.barrier()
    tdist.destroy_process_group()


PEER = {}                                                       # default: own all-reduce kernel, NVSwitch multicast when there is one
PEER_P2P = {"DDRL_DP_MULTICAST": "0"}                           # own kernel, plain peer loads / stores
NCCL = {"DDRL_DP_COLLECTIVE": "nccl"}
NCCL_OVERLAP = {"DDRL_DP_COLLECTIVE": "nccl", "DDRL_DP_OVERLAP": "1"}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind,B,env,iters", [("pong", 16, PEER, 2), ("pong", 16, PEER_P2P, 2), ("pong", 16, NCCL, 2), ("pong", 16, NCCL_OVERLAP, 2),
                                              ("pong", 37, NCCL_OVERLAP, 4), ("navimg", 11, PEER, 3), ("navimg", 11, NCCL_OVERLAP, 3),
                                              ("navlaser", 6, PEER, 3), ("navlaser", 6, NCCL_OVERLAP, 3)])
def test_two_gpu_learn_equals_single_gpu(kind, B, env, iters):
    """The sharded learner against the single-GPU full batch, for every way the gradient sum can travel: this library's
    peer-memory all-reduce kernel (multicast and plain peer path), NCCL after the pass, and NCCL under the pass (iterations
    2.. run the backward as a chain of segments and reduce each segment's gradient ranges while the next one computes)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q, kind, env, iters)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    r0 = [g for g in got if g[0] != "rank1"][0]
    r1 = [g for g in got if g[0] == "rank1"][0]
    logs, rlogs, flat0, flat_ref, used = r0
    if env.get("DDRL_DP_COLLECTIVE") == "nccl":
        assert used == "nccl"
    elif env.get("DDRL_DP_MULTICAST") == "0":
        assert used in ("peer", "nccl")
    print("collective used:", used)
    assert len(logs) == iters
    for i, (l, r) in enumerate(zip(logs, rlogs)):
        tol = 2e-5 if i < 2 else 2e-3         # later iterations inherit Adam's sign-like first steps on noise-level gradients
        for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss"):
            assert abs(l[k] - r[k]) <= tol * max(1.0, abs(r[k])), (i, k, l[k], r[k])
    assert np.array_equal(flat0, r1[1])                       # replicas stay bit-identical
    d = np.abs(flat0 - flat_ref)
    assert (d > 3e-4 * iters).mean() < 1e-3                   # Adam step-1 sign flips on noise-level grads only


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_net_on_second_device_while_first_is_current():
    """The engine allocates and launches on the NET's device, whatever device the caller left current."""
    from ddrl4nav_b200.runner import make_net
    spec = R.SPECS["navimg"]
    params = R.init_params(spec, seed=3)
    states = R.synth_states("navimg", 9, seed=4)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=5)
    outs = []
    torch.cuda.set_device(0)
    for d in (0, 1):
        dev = torch.device("cuda", d)
        net = make_net("navimg", device=None)
        net.load_state_dict(params)
        net = net.to(dev)
        assert torch.cuda.current_device() == 0
        ds = [s.to(dev) for s in states]
        acts, logp, vals = net.act(ds, play_mode=True)
        net.backward_only(ds, adv.to(dev), a.to(dev), old.to(dev), ret.to(dev))
        net.optimizer_step()
        torch.cuda.synchronize(dev)
        outs.append((acts.cpu(), vals.cpu(), net.flat_grads().cpu(), net._flat.detach().cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-7)
    g0, g1 = outs[0][2], outs[1][2]
    assert float((g0 - g1).abs().max()) <= 5e-6 * float(g0.abs().max())


def _ar_worker(rank, world, port, q, multicast):
    import ctypes as C
    import os
    import torch.distributed as tdist
    import torch.distributed._symmetric_memory as symm_mem
    from ddrl4nav_b200 import _lib
    from ddrl4nav_b200._lib import check, current_stream
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tdist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    lib = _lib.load()
    out = []
    for n in (8, 4096, 3368008, 12800012 // 4 * 4):
        g = symm_mem.empty(n, dtype=torch.float32, device=dev)
        f = symm_mem.empty(lib.ddrl_peer_allreduce_flag_bytes() // 4, dtype=torch.int32, device=dev)
        f.zero_()
        hg, hf = symm_mem.rendezvous(g, tdist.group.WORLD), symm_mem.rendezvous(f, tdist.group.WORLD)
        torch.cuda.synchronize()
        tdist.barrier()
        bufs = (C.c_void_p * world)(*[int(p) for p in hg.buffer_ptrs])
        flags = (C.c_void_p * world)(*[int(p) for p in hf.buffer_ptrs])
        mc = C.c_void_p(int(hg.multicast_ptr)) if (multicast and int(hg.multicast_ptr or 0)) else None
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        for seq in range(1, 4):                                       # the flag slots are reused call after call
            x = torch.randn(n, device=dev, generator=gen)
            ref = x.clone()
            tdist.all_reduce(ref)
            g.copy_(x)
            check(lib.ddrl_peer_allreduce_f32(bufs, mc, flags, rank, world, 0, n, seq, current_stream()), "peer_allreduce")
            out.append((n, bool(torch.equal(g, ref)), float((g - ref).abs().max()), mc is not None))
        # a sub-range leaves the rest of the buffer alone
        x = torch.randn(n, device=dev, generator=gen)
        g.copy_(x)
        if n >= 4096:
            check(lib.ddrl_peer_allreduce_f32(bufs, mc, flags, rank, world, 1024, 2048, 4, current_stream()), "peer_allreduce")
            ref = x.clone()
            part = x[1024:3072].clone()
            tdist.all_reduce(part)
            ref[1024:3072] = part
            out.append((n, bool(torch.equal(g, ref)), float((g - ref).abs().max()), mc is not None))
        torch.cuda.synchronize()
        tdist.barrier()
    q.put((rank, out))
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", [True, False])
def test_peer_allreduce_kernel_equals_nccl(multicast):
    """ddrl_peer_allreduce_f32 through the C ABI on symmetric-memory buffers of two ranks: with two ranks a + b has one
    rounding, so the result must equal NCCL's bit for bit -- through the NVSwitch multicast address and through plain peer
    pointers, for tiny, odd-sliced and gradient-sized buffers, with the flag slots reused across calls."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q, multicast)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, out in got:
        assert out and all(o[1] for o in out), (rank, out)
