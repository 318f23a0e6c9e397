"""World-size-2 gloo tests (CPU) of the data-parallel host logic in ddrl4nav_b200/dist.py: row sharding,
the 1/B_global scaling convention, the flat-gradient (+loss tail) all-reduce and the max-over-ranks timing
rule.  The per-rank gradients come from the oracle (CPU autograd), so no GPU is needed; the same calls run
over NCCL on the GPU box (tests/test_gpu_dist.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import restate as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    import torch.distributed as tdist
    from ddrl4nav_b200 import dist
    torch.set_num_threads(2)
    tdist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    spec = R.SPECS["navimg"]
    params = R.init_params(spec, seed=3)
    states = R.synth_states("navimg", B, seed=4)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=5)
    lo, hi = dist.shard_rows(B, rank, world)
    b_global = dist.global_rows(hi - lo, "cpu")
    assert b_global == B
    # local loss = sum over local rows / B_global  ==  (local mean) * (b_local / B_global)
    st = R.LearnState(spec, params)
    sl = lambda t: t[lo:hi]
    losses, raw, _ = R.learn_iteration(st, [sl(s) for s in states], sl(adv), sl(a), sl(old), sl(ret), R.PPOHyper(),
                                       apply_update=False)
    w = (hi - lo) / B
    names = [n for n, _ in R.param_table(spec)]
    P = sum(raw[n].numel() for n in names)
    flat = torch.zeros(P + 8)
    flat[:P] = torch.cat([raw[n].flatten() for n in names]) * w
    flat[P:P + 3] = torch.tensor([losses["ActorLoss"], losses["VLoss"], losses["EntLoss"]]) * w
    dist.allreduce_grads(flat, P)
    t = dist.max_over_ranks(1.0 + rank, "cpu")
    assert t == float(world)
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    tdist.barrier()
    tdist.destroy_process_group()


def test_shard_rows_partition():
    from ddrl4nav_b200.dist import shard_rows
    for n in (0, 1, 7, 8, 1024, 65536):
        for w in (1, 2, 3, 8):
            parts = [shard_rows(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    B, world = 6, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    flat = np.load(os.path.join(str(tmp_path), "flat.npy"))
    spec = R.SPECS["navimg"]
    params = R.init_params(spec, seed=3)
    states = R.synth_states("navimg", B, seed=4)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=5)
    st = R.LearnState(spec, params)
    losses, raw, _ = R.learn_iteration(st, states, adv, a, old, ret, R.PPOHyper(), apply_update=False)
    names = [n for n, _ in R.param_table(spec)]
    full = torch.cat([raw[n].flatten() for n in names]).numpy()
    P = full.size
    assert np.allclose(flat[:P], full, rtol=1e-4, atol=1e-6 * np.abs(full).max())
    assert np.allclose(flat[P:P + 3], [losses["ActorLoss"], losses["VLoss"], losses["EntLoss"]], rtol=1e-5, atol=1e-6)
