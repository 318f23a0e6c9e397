"""Which parameter tensors differ between repeated backward passes?  usage: dbg_det.py kind B det(0|1) passes [predefault]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate as R
from ddrl4nav_b200 import kernels
from ddrl4nav_b200.runner import make_net
kind, B, det, passes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
pre = len(sys.argv) > 5
spec = R.SPECS[kind]
params = R.init_params(spec, seed=11)
states = R.synth_states(kind, B, seed=9)
a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=9)
net = make_net(kind, device=None, gemm_mode="tc3")
net.load_state_dict(params)
net = net.to("cuda")
ds = [s.cuda() for s in states]
dv = [t.cuda() for t in (adv, a, old, ret)]
if pre:
    net.backward_only(ds, dv[0], dv[1], dv[2], dv[3])
kernels.set_deterministic(bool(det))
grads = []
for i in range(passes):
    net.backward_only(ds, dv[0], dv[1], dv[2], dv[3], obs_unchanged=i > 0)
    grads.append(net._grads.clone())
names = [n for n, _ in net.named_parameters()]
sizes = [p.numel() for _, p in net.named_parameters()]
for i in range(1, passes):
    bad = []
    for n, o, sz in zip(names, net._offsets, sizes):
        x, y = grads[0][o:o + sz], grads[i][o:o + sz]
        if not torch.equal(x, y):
            bad.append((n, int((x != y).sum()), "%.1e" % (float((x - y).abs().max()) / float(x.abs().max()))))
    print("pass", i, "vs 0:", bad if det else [b for b in bad if float(b[2]) > 2e-5])
# pattern of the large differences in the first conv of the ped-map tower (flat index = n * 147 + k)
for i in range(1, passes):
    for n, o, sz in zip(names, net._offsets, sizes):
        if not n.endswith("pre.conv1.weight"):
            continue
        x, y = grads[0][o:o + sz].view(64, -1), grads[i][o:o + sz].view(64, -1)
        d = (x - y).abs()
        big = d > 1e-4 * x.abs().max()
        if big.any():
            ks = sorted(set(big.nonzero()[:, 1].tolist()))
            ns = sorted(set(big.nonzero()[:, 0].tolist()))
            print("pass", i, n, "big diffs:", int(big.sum()), "k:", ks[:40], "n:", ns[:70])
