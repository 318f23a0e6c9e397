import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate as R
from ddrl4nav_b200.runner import make_net
kind = sys.argv[1] if len(sys.argv) > 1 else "pong"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
spec = R.SPECS[kind]
params = R.init_params(spec, seed=11)
states = R.synth_states(kind, B, seed=5)
a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=9)
res = {}
for mode in ("simt", "tc", "tc2"):
    net = make_net(kind, device=None, gemm_mode=mode)
    net.load_state_dict(params, strict=True)
    net = net.to(dev)
    net.backward_only([s.to(dev) for s in states], adv.to(dev), a.to(dev), old.to(dev), ret.to(dev))
    res[mode] = {k: v.clone().cpu().double() for k, v in net.named_grads().items()}
for n in res["simt"]:
    ref = res["simt"][n]
    sc = max(float(ref.abs().max()), 1e-30)
    print("%-34s tc %.2e  tc2 %.2e" % (n, float((res["tc"][n] - ref).abs().max()) / sc, float((res["tc2"][n] - ref).abs().max()) / sc))
