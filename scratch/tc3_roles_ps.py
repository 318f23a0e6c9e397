#!/usr/bin/env python
"""Role-level accounting of the PRE-SPLIT launches (Pong first conv: forward + weight gradient) inside a real learn step:
   make -C ddrl4nav_b200/csrc timing3ps
   DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing_ps.so python scratch/tc3_roles_ps.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ddrl4nav_b200 import _lib
from ddrl4nav_b200.data import Experience
from ddrl4nav_b200.runner import make_net

lib = _lib.load()
lib.ddrl_tc3_timing_read.restype = C.c_int
lib.ddrl_tc3_timing_read.argtypes = [C.c_void_p, C.c_int]
NAMES = {0: ("producer", ["empty"]), 1: ("mma-chunk", ["mfree", "full/aready", "issue+commit"]), 2: ("mma-corr", ["cfree", "full/aready", "issue+commit"]),
         4: ("epilogue", ["mfull", "cfull", "stores", "of which read-out wait + barrier", "...+ compute + panel writes", "proxy fence", "panel barrier"]),
         5: ("w-producer", ["empty"]), 6: ("w-mma-chunk", ["mfree", "aready", "issue+commit"]),
         8: ("w-epilogue", ["mfull", "cfull", "atomics", "full+afree"])}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
net = make_net("pong", device=dev, gemm_mode="tc3", TRAINING_ITER_TIME=10)
states_h, adv_h, ret_h = bench.synth_batch_host("pong", B, seed=100)
states_d = [s.to(dev) for s in states_h]
acts, logp, _ = net.act(states_d)
old = logp + 0.15 * torch.randn(B, device=dev)
exp = Experience(states=states_d, advs=adv_h.to(dev), actions=acts, old_logps=old, values=ret_h.to(dev)[None])
for _ in range(2):
    for _x in net.learn(exp):
        pass
torch.cuda.synchronize()
lib.ddrl_tc3_timing_read(None, 1)
for _x in net.learn(exp):
    pass
torch.cuda.synchronize()
buf = (C.c_ulonglong * 96)()
lib.ddrl_tc3_timing_read(buf, 1)
for role, (nm, ws) in NAMES.items():
    life = buf[role * 8 + 7] or 1
    print("   %-12s lifetime %8.0f kcyc/CTA-launch | " % (nm, life / 1e3 / 10 / 148) +
          "  ".join("%s %4.1f%%" % (w, 100.0 * buf[role * 8 + k] / life) for k, w in enumerate(ws)))
