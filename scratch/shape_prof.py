#!/usr/bin/env python
"""Per-launch-shape device time of ONE learn step (10 PPO iterations): DDRL_PROF_SHAPES=1 tags every GEMM/conv
launch with its shape.  usage: python scratch/shape_prof.py [pong|navlaser|navimg] [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DDRL_PROF_SHAPES"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from ddrl4nav_b200 import _lib  # noqa: E402
from ddrl4nav_b200.data import Experience  # noqa: E402
from ddrl4nav_b200.runner import make_net  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "pong"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[kind]["batch"]
dev = torch.device("cuda", 0)
lib = _lib.load()
net = make_net(kind, device=dev, gemm_mode=os.environ.get("DDRL_GEMM_MODE", "tc3"), TRAINING_ITER_TIME=10)
states_h, adv_h, ret_h = bench.synth_batch_host(kind, B, seed=100)
states_d = [s.to(dev) for s in states_h]
acts, logp, _ = net.act(states_d)
old = logp + 0.15 * torch.randn(B, device=dev)
exp = Experience(states=states_d, advs=adv_h.to(dev), actions=acts, old_logps=old, values=ret_h.to(dev)[None])
for _ in range(2):
    for _x in net.learn(exp):
        pass
torch.cuda.synchronize()
lib.ddrl_prof_start(C.c_void_p(torch.cuda.current_stream().cuda_stream))
for _x in net.learn(exp):
    pass
buf = C.create_string_buffer(1 << 18)
lib.ddrl_prof_stop(buf, len(buf))
rows = []
for line in buf.value.decode().splitlines():
    name, ms, n, work = line.split()
    rows.append((float(ms), name, int(n), float(work)))
tot = sum(r[0] for r in rows)
print("workload %s B=%d: one learn step = %.3f ms (sum of launches)" % (kind, B, tot))
for ms, name, n, work in sorted(rows, reverse=True):
    rate = work / (ms / 1e3) / 1e12 if work else 0.0
    print("%-52s n=%4d  %9.3f ms  %5.1f%%  %8.4f ms/launch  %7.1f T(FLOP|B)/s" % (name, n, ms, 100 * ms / tot, ms / n, rate))
