"""GPU parity tests of the whole actor-learner path (PPO.act / PPO.forward / PPO.learn through the
C ABI) against the oracle restatement and the reference-generated golden vectors.

Tolerance: fp32, allclose(rtol, atol = atol_scale * max|ref| per tensor).  Targets: forward values /
probs 1e-5; gradients 1e-5 of the tensor's max (looser rtol 1e-3 elementwise because near-zero
entries of a gradient carry only rounding noise); losses 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda"
# every engine is checked against the same oracle: CUDA-core fp32, tcgen05 (smem split), tcgen05 (TMEM operand)
GEMM_MODES = os.environ.get("DDRL_TEST_GEMM_MODE", "simt,tc,tc2,tc3").split(",")
GEMM_MODE = GEMM_MODES[0]


@pytest.fixture(autouse=True, params=GEMM_MODES)
def engine(request):
    global GEMM_MODE
    GEMM_MODE = request.param
    yield request.param


def rel_err(a, b):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def make(kind, seed=11, **hyper):
    from ddrl4nav_b200.runner import make_net
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=seed)
    net = make_net(kind, device=None, gemm_mode=GEMM_MODE, **hyper)
    missing = net.load_state_dict(params, strict=True)
    return net.to(DEV), spec, params


@pytest.mark.parametrize("kind,B", [("pong", 8), ("navimg", 6), ("navlaser", 4), ("navped", 5)])
def test_forward_matches_golden_and_oracle(golden_dir, kind, B):
    g = np.load(os.path.join(golden_dir, f"net_{kind}.npz"))
    net, spec, params = make(kind)
    assert [n for n, _ in net.named_parameters()] == list(g["names"])
    states = R.synth_states(kind, B, seed=5)
    dstates = [s.to(DEV) for s in states]
    act = torch.from_numpy(g["act"])
    (pi, logp), values = net(dstates, act.to(DEV))
    assert values[0].shape == (B, 1)
    assert rel_err(torch.stack(values, 0), g["values"]) < 1e-5
    if spec.dist == "categorical":
        assert rel_err(pi.probs, g["probs"]) < 1e-5
        (praw, _), _ = net(dstates, play_mode=True)
        assert rel_err(praw, g["probs_raw"]) < 1e-5
        assert rel_err(pi.entropy(), g["entropy"]) < 1e-5
    else:
        assert rel_err(pi.loc, g["mu"]) < 1e-5
        assert rel_err(pi.scale, g["std"]) < 1e-6
    assert rel_err(logp, g["logp"]) < 2e-5


@pytest.mark.parametrize("kind,B", [("pong", 33), ("navimg", 9), ("navlaser", 5), ("pong", 1), ("navped", 7)])
def test_act_matches_oracle(kind, B):
    net, spec, params = make(kind)
    states = R.synth_states(kind, B, seed=21)
    gen = torch.Generator().manual_seed(4)
    draw = torch.rand(B, generator=gen) if spec.dist == "categorical" else torch.randn(B, spec.act_dim, generator=gen)
    ra, rlp, rv = R.forward_body(spec, params, states, draw)
    a, lp, v, pi = net.act([s.to(DEV) for s in states], draw=draw.to(DEV), want_pi=True)
    assert v.shape == (1, B, 1) and rel_err(v, rv) < 1e-5
    if spec.dist == "categorical":
        # bit-exact given the kernel's probs; vs the oracle's probs at most a rare boundary flip
        own = R.sample_categorical(pi.cpu().numpy(), draw.numpy())
        assert np.array_equal(a.cpu().numpy().astype(np.int64), own)
        same = (a.cpu() == ra)
        assert same.float().mean() >= 1 - 2.0 / max(B, 1) - 1e-9
        assert rel_err(lp.cpu()[same], rlp[same]) < 2e-5
    else:
        assert rel_err(a, ra) < 1e-5 and rel_err(lp, rlp) < 2e-5
    # play mode
    pa, plp, pv = R.forward_body(spec, params, states, None, play_mode=True)
    a, lp, v = net.act([s.to(DEV) for s in states], play_mode=True)
    assert float(lp.abs().max()) == 0.0
    if spec.dist == "categorical":
        assert (a.cpu() == pa).float().mean() >= 1 - 1.0 / max(B, 1) - 1e-9
    else:
        assert rel_err(a, pa) < 1e-5


def _learn_case(kind, B, seed=9):
    spec = R.SPECS[kind]
    params = R.init_params(spec, seed=11)
    states = R.synth_states(kind, B, seed=5)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=seed)
    return spec, params, states, a, old, adv, ret


def _oracle_grads(spec, params, states, adv, a, old, ret, flips=None):
    """Oracle gradients; `flips` = {layer: [flat indices]} re-runs the backward with those activation branches flipped."""
    rec = R.TieRecorder(flips)
    R.TIE_HOOK = rec
    try:
        st = R.LearnState(spec, params)
        losses, raw, norm = R.learn_iteration(st, states, adv, a, old, ret, R.PPOHyper(), apply_update=False)
    finally:
        R.TIE_HOOK = None
    return losses, raw, rec.z


TIE_EPS = 1e-6      # an activation with |z| < TIE_EPS * max|z| of its layer is zero to within fp32 rounding: a tie


@pytest.mark.parametrize("kind,B", [("pong", 8), ("navimg", 6), ("navlaser", 4), ("pong", 40), ("navped", 5), ("navlaser3", 4)])
def test_backward_grads_match_oracle(kind, B):
    """Every parameter gradient within 2e-5 of its tensor's max.  relu / leaky_relu are not differentiable at 0, and a
    pre-activation that is zero to fp32 rounding (|z| < 1e-6 max|z|) may take either branch depending on summation
    order (the reference itself differs between its CPU and CUDA paths there).  If the plain comparison fails, the
    residual must be explained EXACTLY by flipping the branch of some of those tie elements in the oracle."""
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    losses, raw, zs = _oracle_grads(spec, params, states, adv, a, old, ret)
    net, _, _ = make(kind)
    net.backward_only([s.to(DEV) for s in states], adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    grads = {n: g.detach().cpu().double() for n, g in net.named_grads().items()}
    scale = {n: max(float(g.abs().max()), 1e-30) for n, g in raw.items()}

    def worst_err(ref):
        return max(float((grads[n] - ref[n].double()).abs().max()) / scale[n] for n in ref)

    worst = worst_err(raw)
    if worst >= 2e-5:
        ties = [(name, int(i)) for name, z in zs.items()
                for i in (z.reshape(-1).abs() < TIE_EPS * float(z.abs().max())).nonzero().flatten().tolist()]
        assert 0 < len(ties) <= 64, (worst, len(ties))
        chosen = {}
        for name, i in ties:
            # flipping tie t changes the gradients by D_t; it is part of the explanation iff the residual projects onto D_t
            _, g1, _ = _oracle_grads(spec, params, states, adv, a, old, ret, {name: [i]})
            num = sum(float(((grads[n] - raw[n].double()) * (g1[n] - raw[n]).double()).sum()) / scale[n] ** 2 for n in raw)
            den = sum(float(((g1[n] - raw[n]).double() ** 2).sum()) / scale[n] ** 2 for n in raw)
            if den > 0 and abs(num / den - 1.0) < 0.2:
                chosen.setdefault(name, []).append(i)
        assert chosen, ("gradient mismatch not explained by activation ties", worst, ties)
        _, raw2, _ = _oracle_grads(spec, params, states, adv, a, old, ret, chosen)
        worst = worst_err(raw2)
        print(kind, B, "activation ties taken on the other branch:", chosen)
    assert worst < 2e-5, worst
    sums = net._grads[net._P:net._P + 3].cpu().numpy()
    assert np.allclose(sums, [losses["ActorLoss"], losses["VLoss"], losses["EntLoss"]], rtol=1e-5, atol=1e-6)
    print(kind, B, "worst grad err / max|g| = %.2e" % worst)


@pytest.mark.parametrize("kind,B", [("pong", 8), ("navimg", 6), ("navlaser", 4), ("navped", 5)])
def test_learn_matches_golden(golden_dir, kind, B):
    g = np.load(os.path.join(golden_dir, f"net_{kind}.npz"))
    from ddrl4nav_b200.data import Experience
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    assert np.array_equal(a.numpy(), g["learn_actions"])
    net, _, _ = make(kind, TRAINING_ITER_TIME=1)
    exp = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                     values=ret.numpy()[None])
    exp.to_tensor(device=DEV)
    out = list(net.learn(exp))
    assert len(out) == 1 and out[0][1] == 1 and out[0][2] is True
    l = out[0][0]
    assert set(l) == {"PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss", "PpoBackUpTime"}
    assert np.allclose([l["PpoTotalLoss"], l["ActorLoss"], l["VLoss"], l["EntLoss"]], g["losses"][0], rtol=1e-5, atol=1e-6)
    # Adam deltas against the reference (step 1 is ~lr*sign(g): allow a tiny fraction of sign flips on noise-level grads)
    sd = net.state_dict()
    for i, n in enumerate(g["names"]):
        d = (sd[n].cpu() - params[n]).flatten()
        d = (d if d.numel() <= 4096 else d[::997]).numpy()
        bad = ~np.isclose(d, g[f"delta_sample_{i}"], rtol=1e-3, atol=2e-7)
        assert bad.mean() < 5e-3, (n, bad.mean())
    # trajectory (loose, chaotic): 3 more iterations
    net2, _, _ = make(kind, TRAINING_ITER_TIME=g["losses_traj"].shape[0])
    traj = [[x["PpoTotalLoss"], x["ActorLoss"], x["VLoss"], x["EntLoss"]] for x, _, _ in net2.learn(exp)]
    assert np.allclose(np.array(traj), g["losses_traj"], rtol=5e-3, atol=5e-4)
    assert net2.update_time == g["losses_traj"].shape[0]


@pytest.mark.parametrize("kind,B", [("navimg", 6), ("pong", 5)])
def test_add_critic_gail_matches_reference_golden(golden_dir, kind, B):
    """V > 1 (nn/ppo.py:63-64,75,95-105): a second critic added the way nn/GAIL.py:116-118 does.  Forward returns both value
    rows; with gail_critic the last critic's loss joins VLoss and its gradient goes where the reference's autograd sends it
    (shared encoder: navimg; the extra critic's own tower: pong).  Golden = the unmodified reference (make_golden.py)."""
    from ddrl4nav_b200.data import Experience
    from ddrl4nav_b200.runner import make_net
    g = np.load(os.path.join(golden_dir, "multicritic.npz"))
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    ret2, extra_params = R.synth_extra_critic(params, ret)
    assert np.array_equal(ret2.numpy(), g[kind + "_ret2"])
    net, _, _ = make(kind, TRAINING_ITER_TIME=1)
    extra = make_net(kind, device=None, gemm_mode=GEMM_MODE).critic          # same class / encoder family as net.critic
    assert [n for n, _ in extra.named_parameters()] == list(g[kind + "_extra_names"]) == list(extra_params)
    extra.load_state_dict(extra_params, strict=True)
    extra = extra.to(DEV)
    net.add_critic(extra)
    net.gail_critic = True
    ds = [s.to(DEV) for s in states]
    (pi, logp), values = net(ds, a.to(DEV))
    assert len(values) == 2 and values[1].shape == (B, 1)
    assert rel_err(torch.stack(values, 0), g[kind + "_values"]) < 1e-5
    _, _, v3 = net.act(ds, play_mode=True)
    assert v3.shape == (2, B, 1) and rel_err(v3, g[kind + "_values"]) < 1e-5
    exp = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                     values=ret2.numpy())
    exp.to_tensor(device=DEV)
    (l, upd, last), = list(net.learn(exp))
    assert np.allclose([l["PpoTotalLoss"], l["ActorLoss"], l["VLoss"], l["EntLoss"]], g[kind + "_losses"], rtol=1e-5, atol=1e-6)
    # PPO's own parameters: the golden holds the CLIPPED gradients (clip_grad_norm_ is in place): scale ours by the same coefficient
    raw = net.named_grads()
    norm = float(torch.sqrt(sum((x.double() ** 2).sum() for x in raw.values())))
    coef = min(1.0, 0.5 / (norm + 1e-6))
    for i, n in enumerate(g[kind + "_names"]):
        ours = (raw[n].flatten() * coef).cpu()
        ours = (ours if ours.numel() <= 4096 else ours[::997]).numpy()
        ref = g[kind + "_grad_sample_%d" % i]
        scale = max(np.sqrt(g[kind + "_grad_digest_%d" % i][2] / max(raw[n].numel(), 1)), 1e-12)     # rms of the tensor
        assert np.abs(ours - ref).max() <= 2e-4 * scale + 2e-5 * np.abs(ref).max(), (n, np.abs(ours - ref).max(), scale)
    # the extra critic's gradients (never clipped: it is not in PPO.parameters())
    for i, (n, p) in enumerate(extra.named_parameters()):
        assert p.grad is not None, n
        ours = p.grad.flatten().cpu()
        ours = (ours if ours.numel() <= 4096 else ours[::997]).numpy()
        ref = g[kind + "_extra_grad_sample_%d" % i]
        scale = max(np.sqrt(g[kind + "_extra_grad_digest_%d" % i][2] / max(p.numel(), 1)), 1e-12)
        assert np.abs(ours - ref).max() <= 2e-4 * scale + 2e-5 * np.abs(ref).max(), (n, np.abs(ours - ref).max(), scale)
    # without gail_critic the extra critic only adds a value row (ppo.py:101-106)
    net2, _, _ = make(kind, TRAINING_ITER_TIME=1)
    net2.add_critic(extra)
    (l2, _, _), = list(net2.learn(exp))
    net3, _, _ = make(kind, TRAINING_ITER_TIME=1)
    exp1 = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                      values=ret.numpy()[None])
    exp1.to_tensor(device=DEV)
    (l3, _, _), = list(net3.learn(exp1))
    assert np.allclose([l2[k] for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss")],
                       [l3[k] for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss")], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind,B", [("pong", 21), ("navimg", 9), ("navlaser", 5)])
def test_graph_replay_matches_stream_launches(monkeypatch, kind, B):
    """Iterations 2..N of a learn call replay two CUDA graphs (backward; optimiser step + weight re-preparation, with the
    step-dependent Adam scalars rewritten per replay).  They must do what the stream launches do (DDRL_NO_GRAPH=1)."""
    from ddrl4nav_b200 import kernels
    from ddrl4nav_b200.data import Experience
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    exp = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                     values=ret.numpy()[None])
    exp.to_tensor(device=DEV)

    def run(no_graph):
        if no_graph:
            monkeypatch.setenv("DDRL_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("DDRL_NO_GRAPH", raising=False)
        net, _, _ = make(kind, TRAINING_ITER_TIME=5)
        ds = [s.to(DEV) for s in states]
        # gradient of the SAME parameters three times: stream launches, graph capture + first launch, graph replay
        gs = []
        for i in range(3):
            net.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV), obs_unchanged=i > 0)
            gs.append(net.flat_grads().clone())
        kernels.launch_count_reset()
        losses = [[l[k] for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss")] for l, _, _ in net.learn(exp)]
        return gs, np.array(losses), net._flat.clone(), net._m.clone(), kernels.launch_count()

    g_graph, l_graph, p_graph, m_graph, n_graph = run(False)
    g_eager, l_eager, p_eager, m_eager, n_eager = run(True)
    # replayed kernel nodes are counted like stream launches; a replayed backward is a chain of segments, each ending with the
    # un-permutation of its own layers' gradients (one launch per segment instead of one per pass): towers x 4 replays more
    assert n_graph == n_eager + (1 if spec.shared else 2) * 4, (n_graph, n_eager)
    for g in g_graph[1:] + g_eager[1:]:
        assert rel_err(g, g_graph[0]) < 5e-6
    assert np.allclose(l_graph[0], l_eager[0], rtol=1e-6, atol=1e-7)
    assert np.allclose(l_graph, l_eager, rtol=2e-3, atol=2e-4), (l_graph, l_eager)
    # Adam's first moment after 5 steps is a smooth function of the gradients (the parameters themselves move by
    # ~lr * sign(g) per step on noise-level gradients): it pins the per-step scalars of the replayed optimiser node
    assert rel_err(m_graph, m_eager) < 2e-3
    assert float((p_graph - p_eager).abs().max()) < 5 * 1e-3 * 5      # <= (steps x largest lr) apart anywhere


@pytest.mark.parametrize("kind,B", [("pong", 21), ("navimg", 9), ("navlaser", 5)])
def test_backward_segments_finalise_their_gradient_ranges(kind, B):
    """ddrl_net_backward_segment (the data-parallel learner's overlap hook): the chain of segments equals the whole pass, and
    the flat-gradient ranges reported for segment k (ddrl_net_tensor_segment) already hold their FINAL values when segment k
    has run -- so a reduction of those ranges may overlap segment k + 1."""
    import ctypes as C
    from ddrl4nav_b200 import _lib
    from ddrl4nav_b200._lib import check, current_stream
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    net, _, _ = make(kind)
    ds = [s.to(DEV) for s in states]
    dv = [t.to(DEV) for t in (adv, a, old, ret)]
    net.backward_only(ds, dv[0], dv[1], dv[2], dv[3])                       # stream launches; stages the observations
    g_whole = net.flat_grads().clone()
    lib = _lib.load()
    args, held, _ = net._bwd_args(ds, dv[0], dv[1], dv[2], dv[3], None, True)
    nseg = C.c_int(0)
    for rep in range(2):                                                      # capture + first launch, then replay
        snaps = []
        k = 0
        while True:
            check(lib.ddrl_net_backward_segment(*args, k, C.byref(nseg), current_stream()), "segment")
            snaps.append(net._grads.clone())
            k += 1
            if k >= nseg.value:
                break
        final = snaps[-1]
        assert rel_err(final[:net._P], g_whole) < 5e-6
        if GEMM_MODE == "simt" and nseg.value == 1:
            continue
        assert nseg.value == (2 if spec.shared else 3), nseg.value
        runs = net._segment_runs(nseg.value)
        covered = torch.zeros(net._P + 4, dtype=torch.bool)
        for k, rk in enumerate(runs):
            for lo, hi in rk:
                assert not covered[lo:hi].any()
                covered[lo:hi] = True
                assert torch.equal(snaps[k][lo:hi], final[lo:hi]), (kind, k, lo, hi)
        assert covered.all()
        # the big linear layers are final before the last segment (that is what makes the overlap worth it)
        early = sum(hi - lo for rk in runs[:-1] for lo, hi in rk)
        assert early > 0.6 * net._P
    # the plain entry point replays the same chain
    net.backward_only(ds, dv[0], dv[1], dv[2], dv[3], obs_unchanged=True)
    assert rel_err(net.flat_grads(), g_whole) < 5e-6


@pytest.mark.parametrize("kind,B", [("pong", 300), ("navimg", 70), ("navlaser", 33)])
def test_deterministic_mode_is_bit_reproducible(kind, B):
    """ddrl_set_deterministic(1): two backward passes on the same inputs give the same BITS (split-K weight gradients, fused
    bias gradients, loss sums and the gradient norm add their block partials in block order), and so do two whole learn()
    trajectories from identical nets; results stay within the usual tolerance of the default (unordered) mode."""
    if GEMM_MODE != "tc3":
        pytest.skip("deterministic mode covers the default (tc3) engine")
    from ddrl4nav_b200 import kernels
    from ddrl4nav_b200.data import Experience
    spec, params, states, a, old, adv, ret = _learn_case(kind, B)
    ds = [s.to(DEV) for s in states]
    dv = [t.to(DEV) for t in (adv, a, old, ret)]
    net, _, _ = make(kind)
    net.backward_only(ds, dv[0], dv[1], dv[2], dv[3])
    g_default = net._grads.clone()
    was = kernels.set_deterministic(True)
    try:
        grads = []
        for i in range(3):                     # stream launches, graph capture + first launch, graph replay
            net.backward_only(ds, dv[0], dv[1], dv[2], dv[3], obs_unchanged=i > 0)
            grads.append(net._grads.clone())
        names = [n for n, _ in net.named_parameters()] + ["<loss sums>"]
        bounds = list(net._offsets) + [net._P, net._P + 8]
        for i in (1, 2):
            bad = [(names[k], int((grads[0][bounds[k]:bounds[k + 1]] != grads[i][bounds[k]:bounds[k + 1]]).sum()),
                    float((grads[0][bounds[k]:bounds[k + 1]] - grads[i][bounds[k]:bounds[k + 1]]).abs().max()),
                    float(grads[0][bounds[k]:bounds[k + 1]].abs().max()))
                   for k in range(len(names)) if not torch.equal(grads[0][bounds[k]:bounds[k + 1]], grads[i][bounds[k]:bounds[k + 1]])]
            assert not bad, (i, bad)
        assert rel_err(grads[0][:net._P], g_default[:net._P]) < 1e-5
        runs = []
        for rep in range(2):
            n2, _, _ = make(kind, TRAINING_ITER_TIME=4)
            exp = Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                             values=ret.numpy()[None])
            exp.to_tensor(device=DEV)
            losses = [[l[k] for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss")] for l, _, _ in n2.learn(exp)]
            runs.append((losses, n2._flat.clone(), n2._m.clone(), n2._v.clone()))
        assert runs[0][0] == runs[1][0]
        for x, y in zip(runs[0][1:], runs[1][1:]):
            assert torch.equal(x, y)
    finally:
        kernels.set_deterministic(was)


def test_micro_batching_equals_single_shot(monkeypatch):
    spec, params, states, a, old, adv, ret = _learn_case("pong", 21)
    net, _, _ = make("pong")
    ds = [s.to(DEV) for s in states]
    net.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    g_full = net.flat_grads().clone()
    monkeypatch.setenv("DDRL_MICRO_BATCH", "4")
    net2, _, _ = make("pong")
    net2.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    assert rel_err(net2.flat_grads(), g_full) < 5e-6
    a1, l1, v1 = net.act(ds, play_mode=True)
    a2, l2, v2 = net2.act(ds, play_mode=True)
    assert torch.equal(a1, a2) and rel_err(v2, v1) < 1e-6


@pytest.mark.parametrize("env", [{"DDRL_S2D": "1"}, {"DDRL_NO_FUSE0": "1"}, {"DDRL_NO_IMPLICIT": "1"}, {"DDRL_NO_S2D_FWD": "1"},
                                 {"DDRL_TC2_ONE_PHASE": "1"}, {"DDRL_NO_PRESPLIT": "1"}, {"DDRL_NO_PRESPLIT_WGRAD": "1"}, {"DDRL_NO_SIGNBITS": "1"},
                                 {"DDRL_TC3_NO_WGRAD_K2": "1"}])
def test_engine_layer_variants_agree(monkeypatch, env):
    """Alternative layer lowerings of the same net (space-to-depth conv1, un-fused first convs, explicit im2col for
    every conv) must give the default lowering's forward values and gradients to fp32 rounding."""
    spec, params, states, a, old, adv, ret = _learn_case("pong", 19)
    ds = [s.to(DEV) for s in states]
    net, _, _ = make("pong")
    net.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    g0 = net.flat_grads().clone()
    _, _, v0 = net.act(ds, play_mode=True)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    net2, _, _ = make("pong")
    net2.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    assert rel_err(net2.flat_grads(), g0) < 1e-5
    _, _, v2 = net2.act(ds, play_mode=True)
    assert rel_err(v2, v0) < 1e-5


def test_presplit_first_conv_equals_in_kernel_split(monkeypatch):
    """tc3 engine, Pong: the staged observation is split ONCE into fp16 hi / lo' planes and the first conv (forward and
    weight gradient) takes its activation tiles straight from TMA (SS MMAs).  Same split values, same K-block order, same
    accumulation chunks as the in-kernel splitter path: forward values must be bit-equal, gradients equal up to the order
    of the split-K atomics."""
    if GEMM_MODE not in (None, "tc3"):
        pytest.skip("pre-split operands are a tc3 lowering")
    spec, params, states, a, old, adv, ret = _learn_case("pong", 37)
    ds = [s.to(DEV) for s in states]
    monkeypatch.setenv("DDRL_PRESPLIT_INFER", "1")        # inference passes take the in-kernel split by default (faster there)
    net, _, _ = make("pong")
    net.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    g1 = net.flat_grads().clone()
    a1, l1, v1 = net.act(ds, play_mode=True)
    monkeypatch.setenv("DDRL_NO_PRESPLIT", "1")
    net2, _, _ = make("pong")
    net2.backward_only(ds, adv.to(DEV), a.to(DEV), old.to(DEV), ret.to(DEV))
    a2, l2, v2 = net2.act(ds, play_mode=True)
    assert torch.equal(v1, v2) and torch.equal(a1, a2) and torch.equal(l1, l2)
    assert rel_err(net2.flat_grads(), g1) < 2e-6


def test_forward_module_streamed_chunks_equal_one_shot():
    """Batches larger than chunk_bytes go through the pinned ring / copy stream in row chunks; same draws -> same
    actions, log-probs and values as the single-shot path (rows are independent)."""
    from ddrl4nav_b200.server import ForwardModule
    net, _, _ = make("pong")
    B = 1100
    states = [s.numpy() for s in R.synth_states("pong", B, seed=5)]
    u = torch.rand(B, generator=torch.Generator().manual_seed(6))
    one = ForwardModule(net, device=DEV, chunk_bytes=1 << 40).step(states, draw=u.to(DEV))
    fm = ForwardModule(net, device=DEV, chunk_bytes=30 << 20, copy_threads=3)      # 256-row chunks, ragged tail
    for _ in range(2):                                                              # second pass re-uses the ring slots
        many = fm.step(states, draw=u.to(DEV))
        assert np.array_equal(one[0], many[0])
        assert np.allclose(one[1], many[1], rtol=1e-6, atol=1e-7) and np.allclose(one[2], many[2], rtol=1e-6, atol=1e-7)
    f64 = fm.step([states[0].astype(np.float64)], draw=u.to(DEV))                   # Pong observations arrive as float64
    assert np.array_equal(one[0], f64[0])


def test_backward_module_prefetch_equals_plain():
    from ddrl4nav_b200.data import Experience
    from ddrl4nav_b200.server import BackwardModule
    spec, params, states, a, old, adv, ret = _learn_case("pong", 24)

    def fresh():
        return Experience(states=[s.numpy() for s in states], advs=adv.numpy(), actions=a.numpy(), old_logps=old.numpy(),
                          values=ret.numpy()[None])
    net1, _, _ = make("pong", TRAINING_ITER_TIME=2)
    logs1 = BackwardModule(net1, device=DEV).train_on(fresh())
    net2, _, _ = make("pong", TRAINING_ITER_TIME=2)
    bm = BackwardModule(net2, device=DEV)
    e = fresh()
    bm.prefetch(e)
    logs2 = bm.train_on(e)
    for l1, l2 in zip(logs1, logs2):
        for k in ("PpoTotalLoss", "ActorLoss", "VLoss", "EntLoss"):
            assert abs(l1[k] - l2[k]) <= 1e-5 * max(1.0, abs(l1[k]))     # split-K atomics: not bit-reproducible run to run


@pytest.mark.parametrize("kind", ["pong", "navimg", "navlaser"])
def test_inference_only_engine_micro_batches(monkeypatch, kind):
    """A net that never trained (the predictor process) runs the inference-only lowering (Pong: space-to-depth conv1,
    no im2col workspace); chunked over micro-batches it must reproduce the single-shot values and greedy actions, and
    both must agree with a net whose workspace is the training one."""
    B = 37
    states = [s.to(DEV) for s in R.synth_states(kind, B, seed=8)]
    net, _, _ = make(kind)
    a0, _, v0 = net.act(states, play_mode=True)
    monkeypatch.setenv("DDRL_MICRO_BATCH", "8")
    net2, _, _ = make(kind)
    a1, _, v1 = net2.act(states, play_mode=True)
    assert torch.equal(a0, a1) and rel_err(v1, v0) < 1e-6
    monkeypatch.delenv("DDRL_MICRO_BATCH")
    net3, _, _ = make(kind)
    acts, logp, _ = net3.act(states)
    net3.backward_only(states, torch.randn(B, device=DEV), acts, logp, torch.randn(B, device=DEV))   # training workspace
    a2, _, v2 = net3.act(states, play_mode=True)
    assert rel_err(v2, v0) < 1e-5
    if kind != "navlaser":
        assert (a2 != a0).float().mean() <= 1.0 / B


def test_shard_sum_equals_full_batch():
    """Data-parallel arithmetic on one GPU: two half-batches scaled by 1/B_global sum to the full-batch gradient."""
    spec, params, states, a, old, adv, ret = _learn_case("pong", 16)
    net, _, _ = make("pong")
    ds = [s.to(DEV) for s in states]
    f = lambda t: t.to(DEV)
    net.backward_only(ds, f(adv), f(a), f(old), f(ret))
    full = net._grads.clone()
    net.backward_only([ds[0][:8]], f(adv[:8]), f(a[:8]), f(old[:8]), f(ret[:8]), b_global=16)
    half = net._grads.clone()
    net.backward_only([ds[0][8:]], f(adv[8:]), f(a[8:]), f(old[8:]), f(ret[8:]), b_global=16)
    half += net._grads
    assert rel_err(half[:net._P], full[:net._P]) < 5e-6
    assert torch.allclose(half[net._P:net._P + 3], full[net._P:net._P + 3], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("kind,B", [("pong", 8192), ("navimg", 2048), ("navlaser", 1024)])
def test_full_size_row_independence_and_shard_linearity(kind, B):
    """BASELINE rows per GPU (too large for the CPU oracle): size-independent properties instead.
    (1) Samples are independent: permuting the rows permutes values / greedy actions BIT-exactly, although every row
        then sits in a different tile, image group and tiling phase of the implicit-GEMM convolutions.
    (2) The data-parallel arithmetic: two ragged shards scaled by 1/B_global sum to the full-batch gradient."""
    net, _, _ = make(kind)
    ds = [s.to(DEV) for s in R.synth_states(kind, B, seed=21)]
    gen = torch.Generator(device=DEV).manual_seed(3)
    perm = torch.randperm(B, device=DEV, generator=gen)
    a1, _, v1 = net.act(ds, play_mode=True)
    a2, _, v2 = net.act([s[perm].contiguous() for s in ds], play_mode=True)
    assert torch.equal(v2.reshape(-1), v1.reshape(-1)[perm])
    assert torch.equal(a2, a1[perm])
    acts, logp, _ = net.act(ds)
    old = logp + 0.15 * torch.randn(B, device=DEV, generator=gen)
    adv = torch.randn(B, device=DEV, generator=gen)
    ret = torch.randn(B, device=DEV, generator=gen)
    net.backward_only(ds, adv, acts, old, ret)
    full = net._grads.clone()
    k = B // 3 + 1                                                    # ragged split: neither shard is a multiple of a tile
    parts = None
    for sl in (slice(0, k), slice(k, B)):
        net.backward_only([s[sl].contiguous() for s in ds], adv[sl], acts[sl].contiguous(), old[sl], ret[sl], b_global=B)
        parts = net._grads.clone() if parts is None else parts + net._grads
    # two summation orders of up to 3.3 M-term fp32 reductions (split-K partials, chunk drains): the gradient tolerance of
    # this suite (2e-5 of the tensor's max) applies; measured 0.6 - 1.2e-5 run to run (atomic order)
    assert rel_err(parts[:net._P], full[:net._P]) < 2e-5
    assert torch.allclose(parts[net._P:net._P + 3], full[net._P:net._P + 3], rtol=1e-4, atol=1e-6)


def test_state_dict_and_model_blob_roundtrip():
    net, spec, params = make("navlaser")
    net._ensure_engine()
    sd = net.state_dict()
    assert list(sd.keys()) == [n for n, _ in R.param_table(spec)]
    for n in sd:
        assert torch.equal(sd[n].cpu(), params[n])
    blob = net.model_bytes()
    import struct
    expect = b"".join(struct.pack(">I", p.dim()) + struct.pack(">%dI" % p.dim(), *p.shape) + p.numpy().tobytes()
                      for p in params.values())        # nn/base.py:38-45 format
    assert blob == expect
    net2, _, _ = make("navlaser", seed=3)
    states = R.synth_states("navlaser", 3, seed=1)
    ds = [s.to(DEV) for s in states]
    before = net2.act(ds, play_mode=True)[2].clone()
    net2.load_model_bytes(blob)                       # updatenn_by_redis path
    after = net2.act(ds, play_mode=True)[2]
    ref = net.act(ds, play_mode=True)[2]
    assert not torch.equal(before, after) and torch.equal(after, ref)


def test_cpu_module_fails_loudly():
    from ddrl4nav_b200 import DDRLError
    from ddrl4nav_b200.runner import make_net
    net = make_net("pong", device=None)
    with pytest.raises(DDRLError):
        net.act([torch.zeros(1, 4, 84, 84)])


def test_gae_dropin_on_experiences():
    from types import SimpleNamespace
    from ddrl4nav_b200.agent import accumulate_rewards
    from ddrl4nav_b200.data import Experience
    rng = np.random.default_rng(0)
    T, V, N = 40, 1, 6
    values = rng.standard_normal((T + 1, V, N)).astype(np.float32)
    rewards = rng.standard_normal((T + 1, V, N)).astype(np.float32)
    dones = (rng.random((T + 1, V, N)) < 0.1).astype(np.uint8)
    gam = np.array([[0.99]], dtype=np.float32)
    exps = [Experience(states=[np.zeros((N, 1))], values=values[t].copy(), dones=dones[t].copy()) for t in range(T + 1)]
    out = accumulate_rewards(SimpleNamespace(discounts=gam, landa=0.95, model_dtype=np.float32), exps, rewards.copy())
    ref_ret, ref_adv = R.gae(values, dones, rewards, gam, 0.95)
    assert len(out) == T
    assert np.allclose(np.stack([e.values for e in out]), ref_ret, rtol=1e-5, atol=1e-5)
    assert np.allclose(np.stack([e.advs for e in out]), ref_adv, rtol=1e-5, atol=1e-5)
    assert accumulate_rewards(SimpleNamespace(), [], rewards) == []
