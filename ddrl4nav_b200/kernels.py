"""Tensor-level wrappers over the C ABI (torch is only used for device memory and streams)."""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import DDRLError, PPOHparams, check, current_stream, ptr, require_cuda


def make_hparams(ppo_clip=0.2, dual_clip=3.0, v_coef=1.0, ent_coef=0.05, max_grad_norm=0.5, clip_grad=True,
                 smooth_l1=False, lr=2e-4, lr_actor=5e-5, lr_critic=1e-3, beta1=0.9, beta2=0.999,
                 adam_eps=1e-8) -> PPOHparams:
    """Defaults = USTC_lab/config/config_nn.py:27-57 and torch.optim.Adam's."""
    return PPOHparams(ppo_clip, dual_clip, v_coef, ent_coef, max_grad_norm, int(bool(clip_grad)), int(bool(smooth_l1)),
                      lr, lr_actor, lr_critic, beta1, beta2, adam_eps)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    require_cuda(t)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def set_deterministic(on: bool = True) -> bool:
    """Run-to-run bit-reproducible gradients / losses on the default engine (ddrl_set_deterministic); returns the previous setting."""
    return bool(_lib.load().ddrl_set_deterministic(1 if on else 0))


def gae(values: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor, gamma: Sequence[float], lam: float,
        algo: int = 0):
    """values [T+1,V,N] f32, rewards [>=T,V,N] f32, dones [>=T,V,N] u8 -> (returns [T,V,N], advs [T,N]).
    Agents._accumulate_rewards (USTC_lab/agent/agent.py:124-140)."""
    lib = _lib.load()
    values = _f32c(values)
    rewards = _f32c(rewards)
    require_cuda(dones)
    if dones.dtype != torch.uint8 or not dones.is_contiguous():
        dones = dones.to(torch.uint8).contiguous()
    Tp1, V, N = values.shape
    T = Tp1 - 1
    assert rewards.shape[0] >= T and dones.shape[0] >= T and tuple(rewards.shape[1:]) == (V, N)
    ret = torch.empty((T, V, N), dtype=torch.float32, device=values.device)
    adv = torch.empty((T, N), dtype=torch.float32, device=values.device)
    gamma = [float(x) for x in gamma]
    if len(gamma) != V:
        raise DDRLError("gae: %d discounts for %d value rows (agent/agent.py:79 builds one per value row)" % (len(gamma), V))
    g = (C.c_float * V)(*gamma)
    check(lib.ddrl_gae_f32(ptr(values), ptr(rewards), ptr(dones), g, float(lam), T, V, N, ptr(ret), ptr(adv), algo,
                           current_stream()), "ddrl_gae_f32")
    return ret, adv


def gae_tempo(values: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor, durations: torch.Tensor, table,
              lam: float, out_f64: bool = True):
    """values [T+1,V,N] f32, rewards/dones [>=T,V,N], durations [>=T] int32, table [<=101] float64 (host)
    -> (returns [T,V,N], advs [T,N]) float64 (or float32).  Agents._accumulate_tempo_rewards (agent/agent.py:142-160)."""
    lib = _lib.load()
    values = _f32c(values)
    rewards = _f32c(rewards)
    require_cuda(dones)
    require_cuda(durations)
    if dones.dtype != torch.uint8 or not dones.is_contiguous():
        dones = dones.to(torch.uint8).contiguous()
    if durations.dtype != torch.int32 or not durations.is_contiguous():
        durations = durations.to(torch.int32).contiguous()
    Tp1, V, N = values.shape
    T = Tp1 - 1
    assert rewards.shape[0] >= T and dones.shape[0] >= T and durations.shape[0] >= T and tuple(rewards.shape[1:]) == (V, N)
    tab = [float(x) for x in table]
    assert 1 <= len(tab) <= 101
    if T > 0:
        # the reference indexes the table with the duration and raises IndexError outside it (agent/agent.py:151)
        lo, hi = int(durations[:T].min()), int(durations[:T].max())
        if lo < 0 or hi >= len(tab):
            raise IndexError("gae_tempo: durations span [%d, %d] but the discount table has %d entries" % (lo, hi, len(tab)))
    dt = torch.float64 if out_f64 else torch.float32
    ret = torch.empty((T, V, N), dtype=dt, device=values.device)
    adv = torch.empty((T, N), dtype=dt, device=values.device)
    check(lib.ddrl_gae_tempo(ptr(values), ptr(rewards), ptr(dones), ptr(durations), (C.c_double * len(tab))(*tab), len(tab),
                             C.c_double(float(lam)), T, V, N, ptr(ret), ptr(adv), int(out_f64), current_stream()),
          "ddrl_gae_tempo")
    return ret, adv


def sample_categorical_probs(probs: torch.Tensor, u: Optional[torch.Tensor]):
    """(probs [B,A], u [B] | None) -> (action [B] f32, logp [B] f32); server/utils.py:20-47."""
    lib = _lib.load()
    probs = _f32c(probs)
    B, A = probs.shape
    action = torch.empty(B, dtype=torch.float32, device=probs.device)
    logp = torch.empty(B, dtype=torch.float32, device=probs.device)
    if u is not None:
        u = _f32c(u)
    check(lib.ddrl_sample_categorical_probs(ptr(probs), A, ptr(u), B, A, ptr(action), ptr(logp), current_stream()),
          "ddrl_sample_categorical_probs")
    return action, logp


def categorical_head(logits: torch.Tensor, u: Optional[torch.Tensor], want_probs: bool = True):
    lib = _lib.load()
    logits = _f32c(logits)
    B, A = logits.shape
    action = torch.empty(B, dtype=torch.float32, device=logits.device)
    logp = torch.empty(B, dtype=torch.float32, device=logits.device)
    probs = torch.empty((B, A), dtype=torch.float32, device=logits.device) if want_probs else None
    if u is not None:
        u = _f32c(u)
    check(lib.ddrl_categorical_head(ptr(logits), A, ptr(u), B, A, ptr(action), ptr(logp), ptr(probs), current_stream()),
          "ddrl_categorical_head")
    return action, logp, probs


def gaussian_head(mu: torch.Tensor, log_std: torch.Tensor, eps: Optional[torch.Tensor]):
    lib = _lib.load()
    mu = _f32c(mu)
    log_std = _f32c(log_std)
    B, A = mu.shape
    action = torch.empty((B, A), dtype=torch.float32, device=mu.device)
    logp = torch.empty(B, dtype=torch.float32, device=mu.device)
    if eps is not None:
        eps = _f32c(eps)
    check(lib.ddrl_gaussian_head(ptr(mu), A, ptr(log_std), ptr(eps), B, A, ptr(action), ptr(logp), current_stream()),
          "ddrl_gaussian_head")
    return action, logp


def ppo_loss_categorical(logits, actions, old_logp, adv, returns, v, hp: PPOHparams, shared: bool,
                         b_global: Optional[int] = None):
    """-> (dlogits [B,A], dv [B], loss_sums [4] = {actor, v, entropy, 0}); nn/ppo.py:85-108."""
    lib = _lib.load()
    logits, actions, old_logp, adv, returns, v = map(_f32c, (logits, actions, old_logp, adv, returns, v))
    B, A = logits.shape
    dlogits = torch.empty_like(logits)
    dv = torch.empty(B, dtype=torch.float32, device=logits.device)
    sums = torch.zeros(4, dtype=torch.float32, device=logits.device)
    check(lib.ddrl_ppo_loss_categorical(ptr(logits), A, ptr(actions), ptr(old_logp), ptr(adv), ptr(returns), ptr(v), B, A,
                                        1.0 / float(b_global or B), C.byref(hp), int(shared), ptr(dlogits), A, ptr(dv),
                                        ptr(sums), current_stream()), "ddrl_ppo_loss_categorical")
    return dlogits, dv, sums


def ppo_loss_gaussian(mu, log_std, actions, old_logp, adv, returns, v, hp: PPOHparams, shared: bool,
                      b_global: Optional[int] = None):
    """-> (dmu [B,A], dv [B], dlog_std [A], loss_sums [4])."""
    lib = _lib.load()
    mu, log_std, actions, old_logp, adv, returns, v = map(_f32c, (mu, log_std, actions, old_logp, adv, returns, v))
    B, A = mu.shape
    dmu = torch.empty_like(mu)
    dv = torch.empty(B, dtype=torch.float32, device=mu.device)
    dls = torch.zeros(A, dtype=torch.float32, device=mu.device)
    sums = torch.zeros(4, dtype=torch.float32, device=mu.device)
    check(lib.ddrl_ppo_loss_gaussian(ptr(mu), A, ptr(log_std), ptr(actions), ptr(old_logp), ptr(adv), ptr(returns), ptr(v),
                                     B, A, 1.0 / float(b_global or B), C.byref(hp), int(shared), ptr(dmu), A, ptr(dv),
                                     ptr(dls), ptr(sums), current_stream()), "ddrl_ppo_loss_gaussian")
    return dmu, dv, dls, sums


def clip_adam(params, grads, m, v, seg_begin: Sequence[int], seg_lr: Sequence[float], step: int, hp: PPOHparams):
    """In-place fused clip_grad_norm_ + Adam over flat fp32 buffers; returns the pre-clip norm (device scalar)."""
    lib = _lib.load()
    require_cuda(params, grads, m, v)
    n = params.numel()
    nseg = len(seg_lr)
    sb = (C.c_int64 * (nseg + 1))(*[int(x) for x in seg_begin])
    sl = (C.c_float * nseg)(*[float(x) for x in seg_lr])
    norm = torch.empty(1, dtype=torch.float32, device=params.device)
    check(lib.ddrl_clip_adam(ptr(params), ptr(grads), ptr(m), ptr(v), n, sb, sl, nseg, int(step), C.byref(hp), ptr(norm),
                             current_stream()), "ddrl_clip_adam")
    return norm


def gemm(form: int, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0,
         mode: str = "simt", out: Optional[torch.Tensor] = None, beta: int = 0):
    """form 0: A[M,K] B[N,K] ; form 1: A[M,K] B[K,N] ; form 2: A[K,M] B[K,N]  ->  C[M,N]."""
    lib = _lib.load()
    A, B = _f32c(A), _f32c(B)
    if form == 0:
        M, K = A.shape; N = B.shape[0]; assert B.shape[1] == K
    elif form == 1:
        M, K = A.shape; N = B.shape[1]; assert B.shape[0] == K
    else:
        K, M = A.shape; N = B.shape[1]; assert B.shape[0] == K
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    if bias is not None:
        bias = _f32c(bias)
    check(lib.ddrl_gemm_f32(_lib.GEMM_MODE[mode], form, M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(out),
                            out.stride(0), ptr(bias), act, beta, current_stream()), "ddrl_gemm_f32")
    return out


def conv_nhwc(op: int, x, w, dy=None, bias=None, stride: int = 1, pad: int = 0, act: int = 0, mask=None, mode: str = "tc"):
    """Implicit-GEMM convolution on NHWC activations (ddrl_conv_nhwc_f32).  x [B,H,W,Cin] (or its shape as a tuple for
    op 1), w [Cout,Cin,KH,KW] reference layout.  op 0: forward -> [B,Ho,Wo,Cout]; op 1: data gradient of dy -> [B,H,W,Cin];
    op 2: weight gradient -> [Cout,Cin,KH,KW]."""
    lib = _lib.load()
    w = _f32c(w)
    Cout, Cin, KH, KW = w.shape
    B, H, W = (x.shape if torch.is_tensor(x) else x)[:3]
    conv1d = H == 1 and KH == 1
    Ho = 1 if conv1d else (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    d = _lib.ConvDesc(B, H, W, Cin, Cout, KH, KW, stride, pad)
    dev = w.device
    if op == 0:
        out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=dev)
    elif op == 1:
        out = torch.zeros((B, H, W, Cin), dtype=torch.float32, device=dev)
    else:
        out = torch.empty_like(w)
    xx = _f32c(x) if torch.is_tensor(x) else None
    dyy = _f32c(dy) if dy is not None else None
    bb = _f32c(bias) if bias is not None else None
    mm = _f32c(mask) if mask is not None else None
    check(lib.ddrl_conv_nhwc_f32(_lib.GEMM_MODE[mode], op, C.byref(d), ptr(xx), ptr(w), ptr(bb), ptr(dyy), act, ptr(mm), ptr(out), current_stream()),
          "ddrl_conv_nhwc_f32")
    return out


def launch_count() -> int:
    return int(_lib.load().ddrl_launch_count())


def launch_count_reset():
    _lib.load().ddrl_launch_count_reset()
