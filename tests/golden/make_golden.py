"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, DRL-Navigation/DDRL4NAV @ 8d52815) in the build container.

    python tests/golden/make_golden.py

The reference pins nothing itself (no tests / seeds / vectors), so these
files are "outputs of the reference itself run here".  Inputs are regenerated from seeds by
``oracle.restate`` (synth_states / init_params) so the fixtures stay small; per-tensor
checksums of the generated parameters are stored to detect RNG-stream drift.
Environment used: torch 2.11.0+cu128 (CPU), numpy 2.3.5, python 3.12.3.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import restate as R  # noqa: E402

SAMPLE_STRIDE = 997


def tensor_digest(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], dtype=np.float64)


def sample_of(t: torch.Tensor) -> np.ndarray:
    f = t.detach().flatten()
    return (f if f.numel() <= 4096 else f[::SAMPLE_STRIDE]).numpy().copy()


def golden_gae():
    ref_shim.import_reference()
    from USTC_lab.agent import Agents
    from USTC_lab.data import Experience
    out = {}
    cases = [("a", 16, 1, 8, 0), ("b", 64, 1, 33, 1), ("c", 37, 2, 5, 2), ("d", 1, 1, 4, 3), ("e", 256, 1, 2, 4)]
    for tag, T, V, N, seed in cases:
        rng = np.random.default_rng(seed)
        values = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        rewards = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        dones = (rng.random((T + 1, V, N)) < 0.1).astype(np.uint8)
        gam = np.array([0.99, 0.999][:V], dtype=np.float32).reshape(V, 1)
        fake_self = SimpleNamespace(discounts=gam, landa=0.95, model_dtype=np.float32)
        exps = [Experience(states=[np.zeros((N, 1))], values=values[t].copy(), dones=dones[t].copy())
                for t in range(T + 1)]
        done = Agents._accumulate_rewards(fake_self, exps, rewards.copy())
        assert len(done) == T
        out[tag + "_values"] = values
        out[tag + "_rewards"] = rewards
        out[tag + "_dones"] = dones
        out[tag + "_gamma"] = gam
        out[tag + "_returns"] = np.stack([e.values for e in done]).astype(np.float32)
        out[tag + "_advs"] = np.stack([e.advs for e in done]).astype(np.float32)
    out["cases"] = np.array([c[0] for c in cases])
    # empty input (agent.py:125-126)
    assert Agents._accumulate_rewards(SimpleNamespace(), [], np.zeros((0,))) == []
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


def golden_gae_tempo():
    """Agents._accumulate_tempo_rewards (agent/agent.py:142-160) run unbound on the unmodified reference."""
    ref_shim.import_reference()
    from USTC_lab.agent import Agents
    from USTC_lab.data import Experience
    out = {}
    cases = [("a", 16, 1, 8, 0), ("b", 67, 1, 33, 1), ("c", 37, 2, 5, 2), ("d", 1, 1, 4, 3)]
    for tag, T, V, N, seed in cases:
        rng = np.random.default_rng(100 + seed)
        values = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        rewards = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        dones = (rng.random((T + 1, V, N)) < 0.1).astype(np.uint8)
        durations = rng.integers(0, 12, size=T + 1).astype(np.int32)
        durations[0] = 100 if T > 1 else durations[0]          # last table entry
        fake_self = SimpleNamespace(tempo_discounts=np.logspace(0, 100, 101, base=0.99), landa=0.95, model_dtype=np.float32)
        exps = [Experience(states=[np.zeros((N, 1))], values=values[t].copy(), dones=dones[t].copy(),
                           rewards=rewards[t].copy(), durations=[int(durations[t])] * 2) for t in range(T + 1)]
        done = Agents._accumulate_tempo_rewards(fake_self, exps)
        assert len(done) == T
        out[tag + "_values"], out[tag + "_rewards"], out[tag + "_dones"] = values, rewards, dones
        out[tag + "_durations"] = durations
        out[tag + "_returns"] = np.stack([e.values for e in done])     # float64 (np.float64 table entry promotes)
        out[tag + "_advs"] = np.stack([e.advs for e in done])
        assert out[tag + "_returns"].dtype == np.float64 and out[tag + "_advs"].dtype == np.float64
    out["cases"] = np.array([c[0] for c in cases])
    assert Agents._accumulate_tempo_rewards(SimpleNamespace(), []) == []
    np.savez_compressed(os.path.join(HERE, "gae_tempo.npz"), **out)


def golden_sampling():
    ref_shim.import_reference()
    import USTC_lab.server.utils as su
    rng = np.random.default_rng(7)
    out = {}
    for tag, B, A in [("a6", 257, 6), ("a28", 129, 28), ("a3", 64, 3)]:
        logits = rng.standard_normal((B, A)).astype(np.float32) * 2
        probs = torch.softmax(torch.from_numpy(logits), dim=-1).numpy()
        if tag == "a3":   # hand-made edge rows
            probs[0] = [0.1, 0.6, 0.3]
            probs[1] = [1.0, 0.0, 0.0]
            probs[2] = [0.0, 0.0, 1.0]
            probs[3] = [0.3, 0.3, 0.4]
        u = rng.random(B).astype(np.float32)
        if tag == "a3":
            u[0] = np.float32(0.7)
            u[1] = np.float32(0.0)
            u[2] = np.float32(0.99999994)
            u[3] = np.float32(1.0)       # >= total -> all False -> index 0
        orig = np.random.rand
        np.random.rand = lambda *shape: u.astype(np.float64)      # feed OUR uniforms into the reference
        try:
            res = su.select_action(probs, PLAY_MODE=False, module_dtype=np.float32)
        finally:
            np.random.rand = orig
        play = su.select_action(probs, PLAY_MODE=True, module_dtype=np.float32)
        out[tag + "_probs"], out[tag + "_u"] = probs, u
        out[tag + "_action"], out[tag + "_logp"] = res[:, 0].copy(), res[:, 1].copy()
        out[tag + "_argmax"] = play[:, 0].copy()
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **out)


def load_into_reference(net, params):
    sd = {k: v.clone() for k, v in params.items()}
    missing = net.load_state_dict(sd, strict=True)
    return missing


def golden_net(kind: str, B: int, iters: int = 4):
    ref_shim.import_reference()
    from USTC_lab.data import Experience
    spec = R.SPECS[kind]
    torch.manual_seed(0)
    net, cfg, cnn = ref_shim.make_ref_net(kind)
    names = [n for n, _ in net.named_parameters()]
    assert names == [n for n, _ in R.param_table(spec)], (names, R.param_table(spec))
    for (n, p), (_, shp) in zip(net.named_parameters(), R.param_table(spec)):
        assert tuple(p.shape) == tuple(shp), (n, p.shape, shp)
    params = R.init_params(spec, seed=11)
    load_into_reference(net, params)
    states = R.synth_states(kind, B, seed=5)
    out = {"param_digest": np.stack([tensor_digest(params[n]) for n in names])}

    # ---- forward (nn/ppo.py:72-75 via server/forward.py:132-146) ----
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        pi, values = net([s.clone() for s in states], play_mode=False)
        dist, _ = pi
        if spec.dist == "categorical":
            out["probs"] = dist.probs.numpy().copy()        # Categorical re-normalised probs
            pi_play, _ = net([s.clone() for s in states], play_mode=True)
            out["probs_raw"] = pi_play[0].numpy().copy()    # softmax output before Categorical()
            act = torch.randint(0, spec.act_dim, (B,), generator=g).float()
        else:
            out["mu"] = dist.loc.numpy().copy()
            out["std"] = dist.scale.numpy().copy()
            act = dist.loc + 0.3 * torch.randn(B, spec.act_dim, generator=g)
        out["act"] = act.numpy().copy()
        out["logp"] = net.actor.log_prob_from_distribution(dist, act).numpy().copy()
        out["entropy"] = dist.entropy().numpy().copy()
        out["values"] = torch.stack(values, dim=0).numpy().copy()    # [V,B,1]

    # ---- learn (nn/ppo.py:77-142) ----
    # (1) ONE iteration from the seeded state: tight check of losses, clipped grads, Adam deltas.
    #     (Adam's first step is ~lr*sign(g), so later iterations amplify 1e-7 noise on near-zero
    #     gradients to 2*lr parameter differences; trajectories are only compared loosely, (2).)
    a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=9)
    out["learn_actions"], out["learn_old"], out["learn_adv"], out["learn_ret"] = \
        a.numpy(), old.numpy(), adv.numpy(), ret.numpy()

    def run(n_iter):
        net_i, _, cnn_i = ref_shim.make_ref_net(kind)
        load_into_reference(net_i, params)
        net_i.training_iter_time = n_iter
        exp = Experience(states=[s.clone() for s in states], advs=adv.clone(), actions=a.clone(),
                         old_logps=old.clone(), values=ret.clone().unsqueeze(0))
        ls = []
        for loss, upd, last in net_i.learn(exp):
            ls.append([loss["PpoTotalLoss"], loss["ActorLoss"], loss["VLoss"], loss["EntLoss"]])
        return net_i, np.array(ls, dtype=np.float64)

    net1, losses1 = run(1)
    out["losses"] = losses1
    after = dict(net1.named_parameters())
    out["delta_digest"] = np.stack([tensor_digest(after[n].detach() - params[n]) for n in names])
    out["grad_digest"] = np.stack([tensor_digest(after[n].grad if after[n].grad is not None
                                                 else torch.zeros_like(after[n])) for n in names])
    out["grad_is_none"] = np.array([after[n].grad is None for n in names])
    for i, n in enumerate(names):
        out["delta_sample_%d" % i] = sample_of(after[n].detach() - params[n])
        gr = after[n].grad if after[n].grad is not None else torch.zeros_like(after[n])
        out["grad_sample_%d" % i] = sample_of(gr)       # CLIPPED grads (clip_grad_norm_ is in place)
    # (2) loss trajectory over `iters` iterations
    _, out["losses_traj"] = run(iters)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "net_%s.npz" % kind), **out)


def golden_easybytes():
    """Wire bytes produced by the reference's own encoder + what its decoder returns for them."""
    ref_shim.import_reference()
    from USTC_lab.data.easybytes import EasyBytes
    eb = EasyBytes("10.0.3.17")
    rng = np.random.default_rng(7)
    out = {}
    # forward states: 3 env processes with 2 / 1 / 3 envs; slots = u8 frames, f64 vector, f16 map, f32 scan
    msgs = b""
    for pid, n in [(0, 2), (5, 1), (131, 3)]:
        frames = rng.integers(0, 256, size=(n, 2, 5, 7), dtype=np.uint8)
        vec = rng.standard_normal((n, 5))                                   # float64, like Pong observations
        half = (rng.random((n, 1, 6, 6)) * 255).astype(np.float16)
        scan = rng.random((n, 1, 33)).astype(np.float32)
        msgs += eb.encode_forward_states(pid, [frames, vec, half, scan])
    ids, slots = eb.decode_forward_states(msgs)
    out["fwd_bytes"] = np.frombuffer(msgs, dtype=np.uint8).copy()
    out["fwd_ids"] = np.array(ids)
    for i, s in enumerate(slots):
        out["fwd_slot_%d" % i] = s
        out["fwd_slot_%d_f32" % i] = torch.tensor(s, dtype=torch.float32).numpy()     # server/forward.py:128-131
    # backward data: [[states...], advs, actions, old_logps, values] + logger dict
    B = 6
    data = [[rng.integers(0, 256, size=(B, 3, 4), dtype=np.uint8), rng.standard_normal((B, 4)).astype(np.float32)],
            rng.standard_normal(B).astype(np.float32), rng.integers(0, 6, size=B).astype(np.float32),
            rng.standard_normal(B).astype(np.float32), rng.standard_normal((1, B)).astype(np.float32)]
    bb = eb.encode_backward_data(data, {"mean_reward": 1.5, "n": 3})
    st, other, logd = eb.decode_backward_data(bb)
    out["bwd_bytes"] = np.frombuffer(bb, dtype=np.uint8).copy()
    for i, s in enumerate(st):
        out["bwd_state_%d" % i] = s
    for i, s in enumerate(other):
        out["bwd_other_%d" % i] = s
    out["bwd_logger_keys"] = np.array(sorted(logd))
    # replies of the Forward thread: [actions, logps, values [V,B,1]] cut per env process (2 / 1 / 3 rows); discrete and
    # 2-d Gaussian actions, float32 (MODULE_NUMPY_DTYPE)
    for tag, act in (("cat", rng.integers(0, 6, size=6).astype(np.float32)), ("gauss", rng.standard_normal((6, 2)).astype(np.float32))):
        arrs = [act, rng.standard_normal(6).astype(np.float32), rng.standard_normal((1, 6, 1)).astype(np.float32)]
        for k, a in enumerate(arrs):
            out["reply_%s_in_%d" % (tag, k)] = a
        for j, b in enumerate(eb.encode_forward_return_data([a.copy() for a in arrs], [2, 1, 3])):
            out["reply_%s_bytes_%d" % (tag, j)] = np.frombuffer(b, dtype=np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, "easybytes.npz"), **out)


def golden_multicritic():
    """add_critic / gail_critic (nn/ppo.py:63-64,75,95-105) on the unmodified reference: a second critic the way
    runner/utils.py:162 makes it (copy.deepcopy(critic), here with re-seeded head weights so its values differ), forward
    values of both critics and ONE learn iteration with gail_critic = True -- shared encoder (navimg: the extra head's
    gradient reaches the shared encoder) and unshared towers (pong: it only reaches the extra critic's own tower)."""
    import copy
    ref_shim.import_reference()
    from USTC_lab.data import Experience
    out = {}
    for kind, B in (("navimg", 6), ("pong", 5)):
        spec = R.SPECS[kind]
        params = R.init_params(spec, seed=11)
        states = R.synth_states(kind, B, seed=5)
        a, old, adv, ret = R.synth_learn_batch(spec, params, states, seed=9)
        ret2, extra_params = R.synth_extra_critic(params, ret)                                # data.values [V=2, B]
        net, cfg, cnn = ref_shim.make_ref_net(kind)
        load_into_reference(net, params)
        extra = copy.deepcopy(net.critic)
        assert [n for n, _ in extra.named_parameters()] == list(extra_params)
        extra.load_state_dict(extra_params, strict=True)
        net.add_critic(extra)
        net.gail_critic = True
        net.training_iter_time = 1
        with torch.no_grad():
            (_, _), values = net([s.clone() for s in states], a.clone())
        exp = Experience(states=[s.clone() for s in states], advs=adv.clone(), actions=a.clone(), old_logps=old.clone(),
                         values=ret2.clone())
        (loss, upd, last), = list(net.learn(exp))
        out[kind + "_ret2"] = ret2.numpy()
        out[kind + "_values"] = torch.stack(values, 0).numpy()                              # [2, B, 1]
        out[kind + "_losses"] = np.array([loss["PpoTotalLoss"], loss["ActorLoss"], loss["VLoss"], loss["EntLoss"]], np.float64)
        names = [n for n, _ in net.named_parameters()]
        after = dict(net.named_parameters())
        out[kind + "_names"] = np.array(names)
        for i, n in enumerate(names):
            gr = after[n].grad if after[n].grad is not None else torch.zeros_like(after[n])
            out[kind + "_grad_sample_%d" % i] = sample_of(gr)                                # CLIPPED grads of PPO's own params
            out[kind + "_grad_digest_%d" % i] = tensor_digest(gr)
        out[kind + "_extra_names"] = np.array(list(extra_params))
        for i, (n, p) in enumerate(extra.named_parameters()):
            out[kind + "_extra_param_digest_%d" % i] = tensor_digest(extra_params[n])
            out[kind + "_extra_grad_sample_%d" % i] = sample_of(p.grad)                      # never clipped (not in PPO.parameters())
            out[kind + "_extra_grad_digest_%d" % i] = tensor_digest(p.grad)
    np.savez_compressed(os.path.join(HERE, "multicritic.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])          # e.g. `make_golden.py navped` regenerates one fixture
    todo = [("gae", golden_gae), ("gae_tempo", golden_gae_tempo), ("sampling", golden_sampling), ("pong", lambda: golden_net("pong", 8)),
            ("navimg", lambda: golden_net("navimg", 6)), ("navlaser", lambda: golden_net("navlaser", 4)),
            ("navped", lambda: golden_net("navped", 5)), ("easybytes", golden_easybytes), ("multicritic", golden_multicritic)]
    for name, fn in todo:
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
