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
        nn.Module.__init__(self)           # not super(): with the reference's class in the MRO its __init__ must not run
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


def _reference_bases(name):
    """When the host application is the reference (its `USTC_lab.nn` is already imported), our actor classes also derive
    from the reference's class of the same name, so that the reference's own `isinstance(net.actor, GaussionActor /
    CategoricalActor)` dispatch (server/forward.py:140-142) holds without patching anything.  Our class comes first in the
    MRO and never calls the reference's __init__ / forward: only the type relation is inherited."""
    import sys
    ref = sys.modules.get("USTC_lab.nn")
    cls = getattr(ref, name, None) if ref is not None else None
    return (Actor, cls) if isinstance(cls, type) and cls is not Actor else (Actor,)


class GaussionActor(*_reference_bases("GaussionActor")):
    DIST = "gaussian"

    def __init__(self, action_output_dim=1, device="cpu", soft_max_grid=True, last_input_dim=512, pre=None,
                 nn_dtype=torch.float32):
        Actor.__init__(self, pre=pre, device=device)
        self.actor_linear = nn.Linear(last_input_dim, action_output_dim)
        self.log_std = nn.Parameter(torch.full((action_output_dim,), -0.5, dtype=nn_dtype))

    def _distribution(self, mu, play_mode=False):
        return mu if play_mode else Normal(mu, torch.exp(self.log_std.detach()))

    def _log_prob_from_distribution(self, pi, act):
        return pi.log_prob(act).sum(axis=-1)


class CategoricalActor(*_reference_bases("CategoricalActor")):
    DIST = "categorical"

    def __init__(self, action_output_dim, device="cpu", soft_max_grid=True, last_input_dim=512, pre=None,
                 nn_dtype=torch.float32):
        Actor.__init__(self, pre=pre, device=device)
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
