import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate as R
from ddrl4nav_b200.runner import make_net
from ddrl4nav_b200 import _lib
lib = _lib.load()
lib.ddrl_net_debug_buffer.restype = C.c_int
lib.ddrl_net_debug_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
kind, B, dev = "pong", 8, "cuda"
spec = R.SPECS[kind]; params = R.init_params(spec, seed=11); states = R.synth_states(kind, B, seed=5)
a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=9)
bufs = {}
sizes = {1: B*400*32, 3: B*81*64, 5: B*49*64, 6: B*49*64, 8: B*81*64, 9: B*400*32}
for mode in ("tc", "tc2"):
    net = make_net(kind, device=None, gemm_mode=mode); net.load_state_dict(params, strict=True); net = net.to(dev)
    net.backward_only([s.to(dev) for s in states], adv.to(dev), a.to(dev), old.to(dev), ret.to(dev))
    for tw in (0, 1):
        for idx, nf in sizes.items():
            out = torch.empty(nf, device=dev)
            rc = lib.ddrl_net_debug_buffer(net._h, tw, idx, out.data_ptr(), nf, None)
            assert rc == 0, rc
            torch.cuda.synchronize()
            bufs[(mode, tw, idx)] = out.cpu().double()
for tw in (0, 1):
    for idx in sizes:
        r, x = bufs[("tc", tw, idx)], bufs[("tc2", tw, idx)]
        e = (r - x).abs()
        sc = float(r.abs().max())
        print("tower", tw, "buf", idx, "max|ref| %.3e" % sc, "err %.2e" % (float(e.max()) / max(sc, 1e-300)), "nbad", int((e > 1e-4 * sc).sum()))
r, x = bufs[("tc", 0, 9)].view(B, 20, 20, 32), bufs[("tc2", 0, 9)].view(B, 20, 20, 32)
bad = ((r - x).abs() > 1e-4 * r.abs().max()).nonzero()
print("bad idx (b,y,x,c) sample:", bad[:12].tolist(), "count", len(bad))
if len(bad):
    import collections
    print("by image", collections.Counter(bad[:, 0].tolist()))
    print("by y parity", collections.Counter((bad[:, 1] % 2).tolist()), "x parity", collections.Counter((bad[:, 2] % 2).tolist()))
    print("by channel", sorted(collections.Counter(bad[:, 3].tolist()).items())[:40])
    b0 = bad[0].tolist(); print("ref", r[tuple(b0)].item(), "got", x[tuple(b0)].item())
