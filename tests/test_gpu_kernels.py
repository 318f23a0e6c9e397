"""GPU parity tests of the stand-alone kernels (K4-K7, GEMM) through the C ABI, against the oracle.

Tolerances (stated once):  integer/index work (sampled actions, argmax)  -> bit-exact;
GAE algo 1 -> bit-exact vs the numpy loop;  everything else fp32 -> allclose(rtol=1e-5, atol=1e-5*max|ref|)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def close(a, b, rtol=1e-5, atol_scale=1e-5):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    ok = np.allclose(a, b, rtol=rtol, atol=atol_scale * scale)
    if not ok:
        err = np.abs(a - b).max() / scale
        print("max err / max|ref| = %.3e" % err)
    return ok


# ------------------------------------------------------------------ K5 GAE
def _gae_inputs(T, V, N, seed, p_done=0.02):
    rng = np.random.default_rng(seed)
    values = rng.standard_normal((T + 1, V, N)).astype(np.float32)
    rewards = rng.standard_normal((T + 1, V, N)).astype(np.float32)
    dones = (rng.random((T + 1, V, N)) < p_done).astype(np.uint8)
    gam = (np.array([0.99, 0.999, 0.9][:V], dtype=np.float32) if V <= 3 else np.linspace(0.9, 0.999, V).astype(np.float32)).reshape(V, 1)
    return values, rewards, dones, gam


def _run_gae(values, rewards, dones, gam, algo):
    from ddrl4nav_b200 import kernels
    ret, adv = kernels.gae(torch.from_numpy(values).to(DEV), torch.from_numpy(rewards).to(DEV),
                           torch.from_numpy(dones).to(DEV), gam.reshape(-1), 0.95, algo)
    return ret.cpu().numpy(), adv.cpu().numpy()


def test_gae_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gae.npz"))
    for tag in g["cases"]:
        v, r, d, gam = g[f"{tag}_values"], g[f"{tag}_rewards"], g[f"{tag}_dones"], g[f"{tag}_gamma"]
        ret, adv = _run_gae(v, r, d, gam, 1)
        assert np.array_equal(ret, g[f"{tag}_returns"]), tag        # bit-exact
        assert np.array_equal(adv, g[f"{tag}_advs"]), tag
        ret4, adv4 = _run_gae(v, r, d, gam, 4)                      # double-buffered sequential schedule: bit-exact too
        assert np.array_equal(ret4, g[f"{tag}_returns"]) and np.array_equal(adv4, g[f"{tag}_advs"]), tag
        for algo in (2, 3, 5) + ((7,) if v.shape[-1] % 4 == 0 else ()):   # time-parallel schedules: carry-in reassociated
            ret2, adv2 = _run_gae(v, r, d, gam, algo)
            assert close(ret2, g[f"{tag}_returns"]) and close(adv2, g[f"{tag}_advs"]), (tag, algo)
        ret0, adv0 = _run_gae(v, r, d, gam, 0)
        assert close(ret0, g[f"{tag}_returns"]) and close(adv0, g[f"{tag}_advs"]), tag


@pytest.mark.parametrize("T,V,N", [(1, 1, 1), (2, 1, 3), (7, 1, 33), (128, 1, 1024), (300, 2, 77), (513, 1, 40), (64, 3, 1000),
                                   (1025, 1, 300), (33, 2, 5000), (16, 1, 9), (17, 8, 3), (260, 2, 4096), (70, 1, 16388)])
def test_gae_vs_oracle(T, V, N):
    v, r, d, gam = _gae_inputs(T, V, N, seed=T * 31 + N)
    ref_ret, ref_adv = R.gae(v, d, r, gam, 0.95)
    for algo in (1, 4):
        ret, adv = _run_gae(v, r, d, gam, algo)
        assert np.array_equal(ret, ref_ret) and np.array_equal(adv, ref_adv), algo
    for algo in (0, 2, 3, 5) + ((7,) if N % 4 == 0 else ()):
        ret, adv = _run_gae(v, r, d, gam, algo)
        assert close(ret, ref_ret) and close(adv, ref_adv), algo


def test_gae_tempo_golden_and_oracle(golden_dir):
    """Agents._accumulate_tempo_rewards: float64 recurrence, per-step discount from the logspace table -- bit-exact."""
    from ddrl4nav_b200 import kernels
    table = np.logspace(0, 100, 101, base=0.99)
    g = np.load(os.path.join(golden_dir, "gae_tempo.npz"))

    def run(v, r, d, du, f64=True):
        T = v.shape[0] - 1
        ret, adv = kernels.gae_tempo(torch.from_numpy(v).to(DEV), torch.from_numpy(r[:max(T, 0)]).to(DEV),
                                     torch.from_numpy(d[:max(T, 0)]).to(DEV), torch.from_numpy(du[:max(T, 0)]).to(DEV),
                                     table, 0.95, out_f64=f64)
        return ret.cpu().numpy(), adv.cpu().numpy()

    for tag in g["cases"]:
        v, r, d, du = g[f"{tag}_values"], g[f"{tag}_rewards"], g[f"{tag}_dones"], g[f"{tag}_durations"]
        ret, adv = run(v, r, d, du)
        assert ret.dtype == np.float64
        assert np.array_equal(ret, g[f"{tag}_returns"]) and np.array_equal(adv, g[f"{tag}_advs"]), tag
        ret32, adv32 = run(v, r, d, du, f64=False)                 # one final rounding (Experience.to_tensor)
        assert np.array_equal(ret32, g[f"{tag}_returns"].astype(np.float32)), tag
        assert np.array_equal(adv32, g[f"{tag}_advs"].astype(np.float32)), tag
    for T, V, N in [(1, 1, 1), (9, 1, 130), (300, 2, 77), (130, 1, 5000), (2050, 1, 64)]:
        rng = np.random.default_rng(T + N)
        v = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        r = rng.standard_normal((T + 1, V, N)).astype(np.float32)
        d = (rng.random((T + 1, V, N)) < 0.05).astype(np.uint8)
        du = rng.integers(0, 101, size=T + 1).astype(np.int32)
        ref_ret, ref_adv = R.gae_tempo(v, d, r, du, 0.99, 0.95)
        ret, adv = run(v, r, d, du)
        assert np.array_equal(ret, ref_ret) and np.array_equal(adv, ref_adv), (T, V, N)


def test_accumulate_tempo_rewards_dropin(golden_dir):
    """agent-level mirror: same signature / side effects as the reference method (agent.py:142-160)."""
    from types import SimpleNamespace
    from ddrl4nav_b200.agent import accumulate_tempo_rewards
    from ddrl4nav_b200.data import Experience
    g = np.load(os.path.join(golden_dir, "gae_tempo.npz"))
    me = SimpleNamespace(tempo_discounts=np.logspace(0, 100, 101, base=0.99), landa=0.95, model_dtype=np.float32)
    assert accumulate_tempo_rewards(me, []) == []
    for tag in g["cases"]:
        v, r, d, du = g[f"{tag}_values"], g[f"{tag}_rewards"], g[f"{tag}_dones"], g[f"{tag}_durations"]
        T = v.shape[0] - 1
        exps = [Experience(states=[np.zeros((v.shape[2], 1))], values=v[t].copy(), dones=d[t].copy(), rewards=r[t].copy(),
                           durations=[int(du[t])]) for t in range(T + 1)]
        out = accumulate_tempo_rewards(me, exps)
        assert len(out) == T
        assert np.array_equal(np.stack([e.values for e in out]), g[f"{tag}_returns"]), tag
        assert np.array_equal(np.stack([e.advs for e in out]), g[f"{tag}_advs"]), tag


def test_gae_empty_and_done_everywhere():
    from ddrl4nav_b200 import kernels
    ret, adv = kernels.gae(torch.zeros(1, 1, 5, device=DEV), torch.zeros(0, 1, 5, device=DEV),
                           torch.zeros(0, 1, 5, dtype=torch.uint8, device=DEV), [0.99], 0.95)
    assert ret.shape == (0, 1, 5) and adv.shape == (0, 5)
    v, r, d, gam = _gae_inputs(50, 1, 64, 3)
    d[:] = 1                                     # every step terminal: adv = r - v exactly
    for algo in (1, 2, 3, 4, 5, 7):
        ret, adv = _run_gae(v, r, d, gam, algo)
        assert np.array_equal(adv, (r[:50, 0] - v[:50, 0]).astype(np.float32) + np.float32(0)) or close(adv, r[:50, 0] - v[:50, 0])


def test_gae_vector_schedule_rejects_ragged_rows():
    """VEC=4 needs N % 4 == 0 (the 4 columns of a thread share one value row): explicit request fails loudly, auto falls back."""
    from ddrl4nav_b200 import kernels
    from ddrl4nav_b200._lib import DDRLError
    v, r, d, gam = _gae_inputs(40, 2, 8191, 5)
    with pytest.raises(DDRLError):
        _run_gae(v, r, d, gam, 7)
    ref_ret, ref_adv = R.gae(v, d, r, gam, 0.95)
    ret, adv = _run_gae(v, r, d, gam, 0)
    assert close(ret, ref_ret) and close(adv, ref_adv)


def test_gae_full_size_properties():
    """BASELINE C3 corner 64k envs x T=2048: algo 1 vs algo 2 agree; power-of-two scaling is exact (linearity)."""
    T, N = 2048, 65536
    gen = torch.Generator(device=DEV).manual_seed(0)
    values = torch.randn(T + 1, 1, N, device=DEV, generator=gen)
    rewards = torch.randn(T, 1, N, device=DEV, generator=gen)
    dones = (torch.rand(T, 1, N, device=DEV, generator=gen) < 0.02).to(torch.uint8)
    from ddrl4nav_b200 import kernels
    ret1, adv1 = kernels.gae(values, rewards, dones, [0.99], 0.95, 1)
    ret2, adv2 = kernels.gae(values, rewards, dones, [0.99], 0.95, 2)
    scale = adv1.abs().max()
    assert float((adv1 - adv2).abs().max() / scale) < 1e-5
    assert float((ret1 - ret2).abs().max() / ret1.abs().max()) < 1e-5
    del ret2, adv2
    for algo in (5, 7):
        ret3, adv3 = kernels.gae(values, rewards, dones, [0.99], 0.95, algo)
        assert float((adv1 - adv3).abs().max() / scale) < 1e-5
        assert float((ret1 - ret3).abs().max() / ret1.abs().max()) < 1e-5
        del ret3, adv3
    ret5, adv5 = kernels.gae(values, rewards, dones, [0.99], 0.95, 4)
    assert torch.equal(adv5, adv1) and torch.equal(ret5, ret1)          # both sequential schedules are bit-identical
    del ret5, adv5
    ret4, adv4 = kernels.gae(values * 4, rewards * 4, dones, [0.99], 0.95, 1)
    assert torch.equal(adv4, adv1 * 4) and torch.equal(ret4, ret1 * 4)
    # ret - adv == values (row 0) up to one rounding
    assert float((ret1[:, 0] - adv1 - values[:T, 0]).abs().max()) < 1e-5 * float(scale)
    # oracle on a column slice
    cols = slice(1000, 1016)
    ref_ret, ref_adv = R.gae(values[:, :, cols].cpu().numpy(), np.concatenate([dones[:, :, cols].cpu().numpy(), np.zeros((1, 1, 16), np.uint8)]),
                             np.concatenate([rewards[:, :, cols].cpu().numpy(), np.zeros((1, 1, 16), np.float32)]),
                             np.array([[0.99]], np.float32), 0.95)
    assert np.array_equal(adv1[:, cols].cpu().numpy(), ref_adv)
    assert np.array_equal(ret1[:, :, cols].cpu().numpy(), ref_ret)


# ------------------------------------------------------------------ K4 sampling / heads
def test_sampling_golden_bit_exact(golden_dir):
    from ddrl4nav_b200 import kernels
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    for tag in ["a6", "a28", "a3"]:
        probs = torch.from_numpy(g[f"{tag}_probs"]).to(DEV)
        u = torch.from_numpy(g[f"{tag}_u"]).to(DEV)
        a, lp = kernels.sample_categorical_probs(probs, u)
        assert np.array_equal(a.cpu().numpy(), g[f"{tag}_action"]), tag
        assert close(lp, g[f"{tag}_logp"], rtol=1e-6, atol_scale=1e-6)
        a, lp = kernels.sample_categorical_probs(probs, None)
        assert np.array_equal(a.cpu().numpy(), g[f"{tag}_argmax"]), tag
        assert float(lp.abs().max()) == 0.0


@pytest.mark.parametrize("B,A", [(1, 2), (1000, 6), (4097, 28), (333, 64)])
def test_sampling_random_bit_exact(B, A):
    from ddrl4nav_b200 import kernels
    from ddrl4nav_b200.server import random_choice_prob_index, select_action
    rng = np.random.default_rng(B + A)
    probs = torch.softmax(torch.from_numpy(rng.standard_normal((B, A)).astype(np.float32) * 3), -1).numpy()
    u = rng.random(B).astype(np.float32)
    ref = R.sample_categorical(probs, u)
    a, _ = kernels.sample_categorical_probs(torch.from_numpy(probs).to(DEV), torch.from_numpy(u).to(DEV))
    assert np.array_equal(a.cpu().numpy().astype(np.int64), ref)
    idx = random_choice_prob_index(torch.from_numpy(probs).to(DEV), u=torch.from_numpy(u).to(DEV))
    assert np.array_equal(idx.cpu().numpy(), ref)
    sel = select_action(torch.from_numpy(probs).to(DEV), u=torch.from_numpy(u).to(DEV), PLAY_MODE=False)
    assert np.array_equal(sel[:, 0].cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(select_action(torch.from_numpy(probs).to(DEV), PLAY_MODE=True)[:, 0].cpu().numpy().astype(np.int64),
                          R.argmax_first(probs))


@pytest.mark.parametrize("B,A", [(513, 6), (100, 28)])
def test_categorical_head(B, A):
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(A)
    logits = torch.randn(B, A, generator=g) * 2
    u = torch.rand(B, generator=g)
    probs = torch.softmax(logits, -1)
    q, _ = R.categorical_normalise(probs)
    a, lp, p_out = kernels.categorical_head(logits.to(DEV), u.to(DEV))
    assert close(p_out, q, rtol=2e-6, atol_scale=1e-6)
    # actions: bit-exact GIVEN the kernel's own probs (the (probs,u) contract); vs torch softmax report mismatch rate
    ref_on_own = R.sample_categorical(p_out.cpu().numpy(), u.numpy())
    assert np.array_equal(a.cpu().numpy().astype(np.int64), ref_on_own)
    mism = (R.sample_categorical(q.numpy(), u.numpy()) != ref_on_own).mean()
    assert mism < 5e-3
    assert close(lp, R.categorical_log_prob(probs, a.cpu()), rtol=1e-5, atol_scale=1e-5)
    a2, lp2, p2 = kernels.categorical_head(logits.to(DEV), None)
    assert np.array_equal(a2.cpu().numpy().astype(np.int64), R.argmax_first(p2.cpu().numpy()))
    assert close(p2, probs, rtol=2e-6, atol_scale=1e-6)


def test_gaussian_head():
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(2)
    B, A = 777, 2
    mu = torch.randn(B, A, generator=g)
    ls = torch.tensor([-0.5, 0.3])
    eps = torch.randn(B, A, generator=g)
    a, lp = kernels.gaussian_head(mu.to(DEV), ls.to(DEV), eps.to(DEV))
    ref_a = R.sample_gaussian(mu, ls, eps)
    assert close(a, ref_a, rtol=1e-6, atol_scale=1e-7)
    assert close(lp, R.gaussian_log_prob(mu, ls, ref_a))
    a, lp = kernels.gaussian_head(mu.to(DEV), ls.to(DEV), None)
    assert torch.equal(a.cpu(), mu) and float(lp.abs().max()) == 0.0


# ------------------------------------------------------------------ K6 loss
def _loss_inputs(B, A, seed, gaussian=False):
    g = torch.Generator().manual_seed(seed)
    head = torch.randn(B, A, generator=g)
    if gaussian:
        ls = torch.tensor([-0.5, 0.1, -1.0, 0.0][:A])
        act = head + torch.exp(ls) * torch.randn(B, A, generator=g)
        logp = R.gaussian_log_prob(head, ls, act)
    else:
        ls = None
        act = torch.randint(0, A, (B,), generator=g).float()
        logp = R.categorical_log_prob(torch.softmax(head, -1), act)
    old = logp + 0.4 * torch.randn(B, generator=g)      # ratios well outside [0.8,1.2] and some > 3
    adv = torch.randn(B, generator=g)
    adv[::17] = 0.0                                      # A == 0 takes the max branch
    ret = torch.randn(B, generator=g)
    v = torch.randn(B, generator=g) * 2
    return head, ls, act, old, adv, ret, v


@pytest.mark.parametrize("A,shared,smooth", [(6, False, False), (28, True, False), (6, True, True), (3, False, True)])
def test_ppo_loss_categorical(A, shared, smooth):
    from ddrl4nav_b200 import kernels
    B = 1500
    logits, _, act, old, adv, ret, v = _loss_inputs(B, A, A)
    spec = R.NetSpec("atari", 4, A, "categorical", shared)
    hp = R.PPOHyper(smooth_l1=smooth)
    x = logits.clone().requires_grad_(True)
    vv = v.clone().requires_grad_(True)
    out = {"probs": torch.softmax(x, -1), "values": vv.unsqueeze(1)}
    out["logp"] = R.categorical_log_prob(out["probs"], act)
    al, vl, ent, total = R.ppo_losses(spec, out, adv, old, ret, hp)
    if shared:
        total.backward()
    else:
        al.backward(retain_graph=True)
        vl.backward()
    khp = kernels.make_hparams(smooth_l1=smooth)
    dl, dv, sums = kernels.ppo_loss_categorical(logits.to(DEV), act.to(DEV), old.to(DEV), adv.to(DEV), ret.to(DEV), v.to(DEV),
                                                khp, shared)
    assert close(sums[:3], [float(al), float(vl), float(ent)], rtol=2e-5, atol_scale=1e-6)
    assert close(dl, x.grad)
    assert close(dv, vv.grad)


@pytest.mark.parametrize("A,shared", [(2, False), (2, True), (4, True)])
def test_ppo_loss_gaussian(A, shared):
    from ddrl4nav_b200 import kernels
    B = 1300
    mu, ls, act, old, adv, ret, v = _loss_inputs(B, A, 10 + A, gaussian=True)
    spec = R.NetSpec("nav1d", 3, A, "gaussian", shared)
    hp = R.PPOHyper()
    x = mu.clone().requires_grad_(True)
    l = ls.clone().requires_grad_(True)
    vv = v.clone().requires_grad_(True)
    out = {"mu": x, "log_std": l, "values": vv.unsqueeze(1), "logp": R.gaussian_log_prob(x, l, act)}
    al, vl, ent, total = R.ppo_losses(spec, out, adv, old, ret, hp)
    if shared:
        total.backward()
    else:
        al.backward(retain_graph=True)
        vl.backward()
    dmu, dv, dls, sums = kernels.ppo_loss_gaussian(mu.to(DEV), ls.to(DEV), act.to(DEV), old.to(DEV), adv.to(DEV), ret.to(DEV),
                                                   v.to(DEV), kernels.make_hparams(), shared)
    assert close(sums[:3], [float(al), float(vl), float(ent)], rtol=2e-5, atol_scale=1e-6)
    assert close(dmu, x.grad) and close(dv, vv.grad)
    assert close(dls, l.grad, rtol=1e-4, atol_scale=1e-5)


# ------------------------------------------------------------------ K7 clip + Adam
@pytest.mark.parametrize("n,nseg", [(1000, 1), (3371847, 2), (5, 2)])
def test_clip_adam_vs_torch(n, nseg):
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    cut = n // 3
    segs = [0, n] if nseg == 1 else [0, cut, n]
    lrs = [2e-4] if nseg == 1 else [5e-5, 1e-3]
    ref = [p0[segs[i]:segs[i + 1]].clone().requires_grad_(True) for i in range(nseg)]
    opts = [torch.optim.Adam([ref[i]], lrs[i]) for i in range(nseg)]
    p = p0.clone().to(DEV)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    hp = kernels.make_hparams()
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (0.01 if step == 2 else 1.0)     # step 2: norm < 0.5 -> coef = 1
        for i in range(nseg):
            ref[i].grad = grad[segs[i]:segs[i + 1]].clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        for o in opts:
            o.step()
        norm = kernels.clip_adam(p, grad.to(DEV), m, v, segs, lrs, step, hp)
        # the kernel accumulates the sum of squares in fp64: tight vs the exact norm; torch's fp32 CPU
        # reduction itself carries ~3e-5 relative error on 3.4 M elements, so only 1e-4 vs it
        assert abs(float(norm) - float(grad.double().norm())) <= 1e-6 * float(total)
        assert abs(float(norm) - float(total)) <= 1e-4 * float(total)
        refp = torch.cat([r.detach() for r in ref])
        d_ref = (refp - p0).double()
        d = (p.cpu() - p0).double()
        # p itself is fp32: the update is only resolved to ~1 ulp(p) = 1.2e-7 * |p|
        tol = 1e-5 * float(d_ref.abs().max()) + 2.4e-7 * (1.0 + p0.abs().double())
        assert bool(((d - d_ref).abs() <= tol).all())


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("mode", ["simt", "tc", "tc2", "tc3"])
@pytest.mark.parametrize("form", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (257, 33, 100), (1000, 512, 3136), (64, 6, 512), (4096, 32, 256), (37, 200, 9),
                                   (128, 32, 32), (300, 64, 1152), (2048, 256, 1600)])
def test_gemm_forms(mode, form, M, N, K):
    from ddrl4nav_b200 import kernels
    if mode in ("tc", "tc2", "tc3"):
        # the tcgen05 paths need 16-byte row strides (TMA) and N >= 16; other shapes stay on the SIMT engine
        lds = {0: (K, K), 1: (K, N), 2: (M, N)}[form]
        if N < 16 or lds[0] % 4 or lds[1] % 4:
            pytest.skip("shape not eligible for the TMA/tcgen05 engine")
    g = torch.Generator().manual_seed(M + N + K + form)
    if form == 0:
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        ref = A.double() @ B.double().T
    elif form == 1:
        A, B = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
        ref = A.double() @ B.double()
    else:
        A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
        ref = A.double().T @ B.double()
    bias = torch.randn(N, generator=g)
    prod = ref
    if mode in ("tc2", "tc3") and form == 2:
        # the TMEM engine's transposed-operand form is the weight gradient: no bias / activation, accumulates
        out = kernels.gemm(form, A.to(DEV), B.to(DEV), None, act=0, mode=mode)
        assert close(out, prod, rtol=1e-5, atol_scale=2e-6)
        out2 = kernels.gemm(form, A.to(DEV), B.to(DEV), None, act=0, mode=mode, out=out.clone(), beta=1)
        assert close(out2, 2 * prod, rtol=1e-5, atol_scale=2e-6)
        return
    out = kernels.gemm(form, A.to(DEV), B.to(DEV), bias.to(DEV), act=2, mode=mode)
    ref = torch.nn.functional.leaky_relu(prod + bias.double(), 0.01)
    assert close(out, ref, rtol=1e-5, atol_scale=2e-6)
    if mode in ("tc2", "tc3"):
        return                       # accumulate-into-C exists only for the weight-gradient form on this engine
    out2 = kernels.gemm(form, A.to(DEV), B.to(DEV), None, act=0, mode=mode, out=out.clone(), beta=1)   # C += A op B
    assert close(out2, ref + prod, rtol=1e-5, atol_scale=2e-6)


@pytest.mark.parametrize("mode", ["simt", "tc", "tc2", "tc3"])
def test_gemm_split_k_wgrad_shape(mode):
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(5)
    K, M, N = 200000, 32, 256
    A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    out = kernels.gemm(2, A.to(DEV), B.to(DEV), mode=mode)
    assert close(out, A.double().T @ B.double(), rtol=1e-5, atol_scale=2e-6)


def test_gemm_tc_is_3xtf32_accurate():
    """The tensor-core engine must be as close to the exact product as the fp32 FFMA engine (not TF32-grade 1e-3)."""
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(9)
    A, B = torch.randn(512, 4096, generator=g), torch.randn(256, 4096, generator=g)
    ref = A.double() @ B.double().T
    e_tc = float((kernels.gemm(0, A.to(DEV), B.to(DEV), mode="tc").cpu().double() - ref).abs().max() / ref.abs().max())
    e_tc2 = float((kernels.gemm(0, A.to(DEV), B.to(DEV), mode="tc2").cpu().double() - ref).abs().max() / ref.abs().max())
    e_tc3 = float((kernels.gemm(0, A.to(DEV), B.to(DEV), mode="tc3").cpu().double() - ref).abs().max() / ref.abs().max())
    e_simt = float((kernels.gemm(0, A.to(DEV), B.to(DEV), mode="simt").cpu().double() - ref).abs().max() / ref.abs().max())
    print("max err / max|ref|: tc %.2e  tc2 %.2e  tc3 %.2e  simt %.2e" % (e_tc, e_tc2, e_tc3, e_simt))
    assert e_tc < 2e-6 and e_tc2 < 2e-6 and e_tc3 < 2e-6 and e_simt < 2e-6


@pytest.mark.parametrize("scale_a,scale_b", [(1.0, 1.0), (1e-7, 3.0), (4e4, 1e-6), (1e-20, 1e-12)])
def test_gemm_tc3_scaled_fp16_split_keeps_fp32_range(scale_a, scale_b):
    """tc3 splits operands into scaled fp16 pairs: tensors far outside the fp16 exponent range, and rows whose magnitudes
    differ by 2^20 inside one tensor, must come out as accurate as the fp32 FFMA engine."""
    from ddrl4nav_b200 import kernels
    g = torch.Generator().manual_seed(13)
    A, B = torch.randn(384, 1024, generator=g) * scale_a, torch.randn(96, 1024, generator=g) * scale_b
    A[::7] *= 2.0 ** -20                                       # small rows next to large ones
    ref = A.double() @ B.double().T
    out = kernels.gemm(0, A.to(DEV), B.to(DEV), mode="tc3").cpu().double()
    assert torch.isfinite(out).all()
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-6
    small = ref[::7]
    assert float((out[::7] - small).abs().max() / small.abs().max()) < 1e-4      # relative to THEIR scale: 2^-20 * 2e-6 would be 0


# ------------------------------------------------------------------ implicit-GEMM convolutions (tap-TMA path)
# (B, H, W, Cin, Cout, KH, KW, stride, pad): the inner conv layers of the three encoders + edge geometries
CONV_CASES = [
    (5, 20, 20, 32, 64, 4, 4, 2, 0),      # Pong conv2   (nn/atari_encoder.py:17)
    (7, 9, 9, 64, 64, 3, 3, 1, 0),        # Pong conv3   (nn/atari_encoder.py:18)
    (3, 24, 24, 64, 128, 3, 3, 1, 1),     # NavPreNet conv2 (nn/nav_encoder.py:18)
    (3, 12, 12, 128, 256, 3, 3, 1, 1),    # NavPreNet conv3
    (2, 22, 22, 64, 128, 5, 5, 1, 1),     # NavPreNet1D conv2 (nn/nav_encoder.py:86)
    (4, 10, 10, 128, 256, 3, 3, 1, 1),    # NavPreNet1D conv3
    (6, 1, 478, 32, 32, 1, 3, 2, 0),      # laser conv1d2 (nn/nav_encoder.py:92)
    (1, 7, 5, 32, 32, 3, 3, 1, 1),        # single image, odd extents
    (130, 6, 6, 32, 96, 3, 3, 2, 1),      # stride 2 with padding, many images per tile, N not a multiple of 64
    (3, 41, 41, 32, 64, 3, 3, 2, 0),      # 20x20 output at stride 2: two pixel-box classes (16x2 + 4x8) in the weight gradient
    (2, 13, 25, 32, 32, 2, 2, 1, 0),      # 12 x 24 output (8x4 + 4x8 boxes would not apply: width 24 = 16 + 8), ragged last boxes
]


def _conv_ref(x_nhwc, w, stride, pad, conv1d):
    x = x_nhwc.permute(0, 3, 1, 2).double()
    if conv1d:
        return torch.nn.functional.conv2d(x, w.double(), None, stride=(1, stride), padding=(0, pad))
    return torch.nn.functional.conv2d(x, w.double(), None, stride=stride, padding=pad)


@pytest.mark.parametrize("mode", ["tc", "tc2", "tc3"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_implicit_forward_dgrad_wgrad(case, mode):
    """ddrl_conv_nhwc_f32 (4-D TMA tap boxes, no im2col) == torch conv2d forward / autograd, fp32 tolerance 1e-5."""
    from ddrl4nav_b200 import kernels
    B, H, W, Cin, Cout, KH, KW, stride, pad = case
    conv1d = H == 1 and KH == 1
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, KH, KW, generator=g) / (Cin * KH * KW) ** 0.5
    bias = torch.randn(Cout, generator=g)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    y_ref = _conv_ref(xr, wr, stride, pad, conv1d)                       # [B, Cout, Ho, Wo]
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy.double())
    # forward (+ bias + leaky relu in the epilogue)
    y = kernels.conv_nhwc(0, x.to(DEV), w.to(DEV), bias=bias.to(DEV), stride=stride, pad=pad, act=2, mode=mode)
    ref = torch.nn.functional.leaky_relu(y_ref.detach() + bias.double()[None, :, None, None], 0.01).permute(0, 2, 3, 1)
    assert close(y, ref, rtol=1e-5, atol_scale=2e-6)
    # data gradient, with the fused activation backward (leaky' of a mask tensor)
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    mask = torch.randn(B, H, W, Cin, generator=g)
    dx = kernels.conv_nhwc(1, (B, H, W), w.to(DEV), dy=dy_nhwc.to(DEV), stride=stride, pad=pad, act=4, mask=mask.to(DEV),
                           mode=mode)
    dx_ref = xr.grad * torch.where(mask > 0, 1.0, 0.01).double()
    assert close(dx, dx_ref, rtol=1e-5, atol_scale=2e-6)
    dx0 = kernels.conv_nhwc(1, (B, H, W), w.to(DEV), dy=dy_nhwc.to(DEV), stride=stride, pad=pad, mode=mode)
    assert close(dx0, xr.grad, rtol=1e-5, atol_scale=2e-6)
    # weight gradient
    dw = kernels.conv_nhwc(2, x.to(DEV), w.to(DEV), dy=dy_nhwc.to(DEV), stride=stride, pad=pad, mode=mode)
    assert close(dw, wr.grad, rtol=1e-5, atol_scale=2e-6)


def test_conv_implicit_rejects_unsupported_channels():
    from ddrl4nav_b200 import kernels
    from ddrl4nav_b200._lib import DDRLError
    x = torch.randn(2, 8, 8, 4, device=DEV)
    w = torch.randn(32, 4, 3, 3, device=DEV)
    with pytest.raises(DDRLError):
        kernels.conv_nhwc(0, x, w)
