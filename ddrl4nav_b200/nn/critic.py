"""Critic head -- USTC_lab/nn/critic.py:6-21 (critic_linear registered BEFORE pre: parameter order)."""
from torch import nn

from .base import PreNet  # noqa: F401


class Critic(nn.Module):
    def __init__(self, device="cpu", last_input_dim=512, pre=None):
        super().__init__()
        self.device = device
        self.critic_linear = nn.Linear(last_input_dim, 1)
        self.pre = pre

    def forward(self, x):
        """Stand-alone call (nn/critic.py:14-21): x = states when the critic owns an encoder, features otherwise; [B,1]."""
        from .. import kernels
        if self.pre is not None:
            x = self.pre(x)
        return kernels.gemm(0, x, self.critic_linear.weight.detach(), self.critic_linear.bias.detach())
