"""Actor heads -- USTC_lab/nn/actor.py:10-101.

`pre` is registered before `actor_linear`, and GaussionActor's `log_std` is a direct Parameter
(so it precedes everything in named_parameters()), exactly as in the reference."""
import torch
from torch import nn
from torch.distributions.categorical import Categorical
from torch.distributions.normal import Normal


class Actor(nn.Module):
    DIST = None

    def __init__(self, **kwargs):
        super().__init__()
        self.pre = kwargs["pre"]
        self.device = kwargs["device"]

    def _distribution(self, x, play_mode=False):
        raise NotImplementedError

    def _log_prob_from_distribution(self, pi, act):
        raise NotImplementedError

    def log_prob_from_distribution(self, pi, act):
        return self._log_prob_from_distribution(pi, act)

    def _head(self, x):
        from .. import kernels
        if self.pre is not None:
            x = self.pre(x)                    # nn/actor.py:28-31: states -> features through the actor's own encoder
        return kernels.gemm(0, x, self.actor_linear.weight.detach(), self.actor_linear.bias.detach())

    def forward(self, x, act=None, play_mode=False):
        """Stand-alone call (nn/actor.py:26-40): x = states when the actor owns an encoder, features [B, last_input_dim]
        otherwise; PPO.forward is the fused path."""
        pi = self._distribution(self._head(x), play_mode)
        log_p = self._log_prob_from_distribution(pi, act) if act is not None else None
        return pi, log_p


class GaussionActor(Actor):
    DIST = "gaussian"

    def __init__(self, action_output_dim=1, device="cpu", soft_max_grid=True, last_input_dim=512, pre=None,
                 nn_dtype=torch.float32):
        super().__init__(pre=pre, device=device)
        self.actor_linear = nn.Linear(last_input_dim, action_output_dim)
        self.log_std = nn.Parameter(torch.full((action_output_dim,), -0.5, dtype=nn_dtype))

    def _distribution(self, mu, play_mode=False):
        return mu if play_mode else Normal(mu, torch.exp(self.log_std.detach()))

    def _log_prob_from_distribution(self, pi, act):
        return pi.log_prob(act).sum(axis=-1)


class CategoricalActor(Actor):
    DIST = "categorical"

    def __init__(self, action_output_dim, device="cpu", soft_max_grid=True, last_input_dim=512, pre=None,
                 nn_dtype=torch.float32):
        super().__init__(pre=pre, device=device)
        self.logits_net = None
        self.action_output_dim = action_output_dim
        self.actor_linear = nn.Linear(last_input_dim, action_output_dim)
        self.soft_max_grid = soft_max_grid

    def _distribution(self, logits, play_mode=False):
        from .. import kernels
        _, _, probs = kernels.categorical_head(logits, None)      # raw softmax (play-mode path)
        return probs if play_mode else Categorical(probs)

    def _log_prob_from_distribution(self, pi, act):
        return pi.log_prob(act)
