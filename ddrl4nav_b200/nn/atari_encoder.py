"""AtariPreNet -- parameter container + engine description (USTC_lab/nn/atari_encoder.py:11-32).

conv(4->32,k8,s4)+leaky -> conv(32->64,k4,s2)+leaky -> conv(64->64,k3,s1)+leaky -> C-major flatten
3136 -> linear 512 (no activation).  Schedule: csrc/net.cu (DDRL_ARCH_ATARI)."""
from torch import nn

from .base import PreNet


class AtariPreNet(PreNet):
    ARCH = "atari"

    def __init__(self, num_inputs=1, last_output_dim=512, device="cpu"):
        super().__init__()
        self.device = device
        chans = [(num_inputs, 32, 8, 4), (32, 64, 4, 2), (64, 64, 3, 1)]
        for i, (cin, cout, k, s) in enumerate(chans, start=1):
            setattr(self, "conv%d" % i, nn.Conv2d(cin, cout, k, stride=s))
        self.linear = nn.Linear(64 * 7 * 7, 512)
        assert last_output_dim == self.linear.out_features

    def engine_in_ch(self):
        return self.conv1.in_channels
