"""Uniform-draw action selection on the device -- USTC_lab/server/utils.py:20-47.

The reference draws ``r = np.random.rand(B)`` inside ``random_choice_prob_index``; here the draw is an
explicit argument (generated with torch on the device when omitted) so results are reproducible and
bit-exact for a given draw."""
from typing import Optional

import torch

from .. import kernels


def random_choice_prob_index(p: torch.Tensor, axis: int = 1, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(p.cumsum(axis) > u).argmax(axis) for p [B, A]; returns int64 indices [B]."""
    assert axis == 1 and p.dim() == 2, "only the [B, A], axis=1 form the reference uses"
    if u is None:
        u = torch.rand(p.shape[0], device=p.device)
    a, _ = kernels.sample_categorical_probs(p, u)
    return a.long()


def select_action(predictions: torch.Tensor, u: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
    """predictions [B, A] -> [B, 2] = (action, old_logp); PLAY_MODE=True -> (argmax, 0)."""
    if kwargs.get("PLAY_MODE", False):
        a, lp = kernels.sample_categorical_probs(predictions, None)
    else:
        if u is None:
            u = torch.rand(predictions.shape[0], device=predictions.device)
        a, lp = kernels.sample_categorical_probs(predictions, u)
    return torch.stack([a, lp], dim=1)
