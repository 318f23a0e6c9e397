"""GAE on the device -- drop-in for ``Agents._accumulate_rewards`` (USTC_lab/agent/agent.py:124-140).

``accumulate_rewards(self, experiences, rewards_step)`` has the reference's exact signature and
side effects (mutates ``.values`` -> returns and ``.advs`` of every experience, returns
``experiences[:-1]``), so it can be bound as a method:  ``Agents._accumulate_rewards = accumulate_rewards``.
``self`` only needs ``discounts`` ([V,1] fp32), ``landa`` and ``model_dtype`` (agent.py:79,117).
"""
from typing import List

import numpy as np
import torch

from .. import kernels


class GAE:
    """Batched form for rollouts that already live on the device: values [T+1,V,N], rewards [>=T,V,N], dones u8."""

    def __init__(self, gamma=(0.99,), lam=0.95, algo=0):
        self.gamma = [float(g) for g in np.asarray(gamma, dtype=np.float32).reshape(-1)]
        self.lam = float(np.float32(lam))      # numpy multiplies the fp32 discounts by a weak python float
        self.algo = algo

    def __call__(self, values: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor):
        return kernels.gae(values, rewards, dones, self.gamma, self.lam, self.algo)


def accumulate_rewards(self, experiences: List, rewards_step: np.ndarray) -> List:
    if len(experiences) == 0:
        return []
    T = len(experiences) - 1
    if T == 0:
        return experiences[:-1]
    dev = torch.device("cuda")
    values = np.stack([np.asarray(e.values, dtype=np.float32) for e in experiences])              # [T+1,V,N]
    dones = np.stack([np.asarray(experiences[t].dones, dtype=np.uint8) for t in range(T)])         # [T,V,N]
    rewards = np.asarray(rewards_step[:T], dtype=np.float32)                                       # [T,R,N]
    V, N = values.shape[1], values.shape[2]
    dones = np.broadcast_to(dones.reshape(T, -1, N), (T, V, N))
    rewards = np.broadcast_to(rewards.reshape(T, -1, N), (T, V, N))
    # schedule 4 (sequential in t, two env columns per thread): bit-exact with the reference's numpy loop -- an actor's rollout
    # is a few hundred steps x a few envs, nothing the time-parallel default (<= 3e-7 relative) would speed up
    gae = GAE(np.asarray(self.discounts, dtype=np.float32).reshape(-1), self.landa, algo=4)
    ret, adv = gae(torch.from_numpy(values).to(dev), torch.from_numpy(np.ascontiguousarray(rewards)).to(dev),
                   torch.from_numpy(np.ascontiguousarray(dones)).to(dev))
    ret, adv = ret.cpu().numpy(), adv.cpu().numpy()
    for t in range(T):
        experiences[t].values = ret[t]
        experiences[t].advs = adv[t]
    return experiences[:-1]


def accumulate_tempo_rewards(self, experiences: List) -> List:
    """Drop-in for ``Agents._accumulate_tempo_rewards`` (USTC_lab/agent/agent.py:142-160): per-step discount
    ``self.tempo_discounts[experiences[t].durations[0]]`` (float64 table), rewards from ``experiences[t].rewards``.
    ``self`` needs ``tempo_discounts`` and ``landa``.  Results are float64 arrays, as in the reference."""
    if len(experiences) == 0:
        return []
    T = len(experiences) - 1
    if T == 0:
        return experiences[:-1]
    dev = torch.device("cuda")
    values = np.stack([np.asarray(e.values, dtype=np.float32) for e in experiences])              # [T+1,V,N]
    V, N = values.shape[1], values.shape[2]
    dones = np.stack([np.asarray(experiences[t].dones, dtype=np.uint8) for t in range(T)]).reshape(T, -1, N)
    rewards = np.stack([np.asarray(experiences[t].rewards, dtype=np.float32) for t in range(T)]).reshape(T, -1, N)
    durations = np.array([int(experiences[t].durations[0]) for t in range(T)], dtype=np.int32)
    dones = np.ascontiguousarray(np.broadcast_to(dones, (T, V, N)))
    rewards = np.ascontiguousarray(np.broadcast_to(rewards, (T, V, N)))
    ret, adv = kernels.gae_tempo(torch.from_numpy(values).to(dev), torch.from_numpy(rewards).to(dev),
                                 torch.from_numpy(dones).to(dev), torch.from_numpy(durations).to(dev),
                                 np.asarray(self.tempo_discounts, dtype=np.float64), float(self.landa), out_f64=True)
    ret, adv = ret.cpu().numpy(), adv.cpu().numpy()
    for t in range(T):
        experiences[t].values = ret[t]
        experiences[t].advs = adv[t]
    return experiences[:-1]
