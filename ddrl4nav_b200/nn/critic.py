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
        """Stand-alone use on FEATURES x [B, last_input_dim] (shared-encoder mode); returns [B,1]."""
        from .. import kernels
        from .._lib import DDRLError
        if self.pre is not None:
            raise DDRLError("Critic with its own encoder runs inside PPO's fused CUDA engine")
        return kernels.gemm(0, x, self.critic_linear.weight.detach(), self.critic_linear.bias.detach())
