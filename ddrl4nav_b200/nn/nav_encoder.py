"""NavPreNet / NavPedPreNet / NavPreNet1D -- parameter containers + engine descriptions
(USTC_lab/nn/nav_encoder.py:12-128).  As in the reference ``last_output_dim`` is ignored (512).
Schedules: csrc/net.cu (DDRL_ARCH_NAV / NAVPED / NAV1D)."""
from torch import nn

from .base import PreNet
from .utils import mlp


def _conv_stack(owner, in_ch, kernels):
    chans = [in_ch, 64, 128, 256]
    for i, k in enumerate(kernels):
        setattr(owner, "conv%d" % (i + 1), nn.Conv2d(chans[i], chans[i + 1], k, stride=1, padding=(1, 1)))


class NavPreNet(PreNet):
    """[conv3x3 p1 + relu + maxpool2] x3 (48->24->12->6) -> 9216 -> fc0+relu -> cat(vec9) -> fc1+relu -> fc2."""
    ARCH = "nav"
    VEC = 9

    def __init__(self, image_channel=1, last_output_dim=512):
        super().__init__()
        _conv_stack(self, image_channel, (3, 3, 3))
        self.fc0 = mlp([(256 * 6 * 6, 512, "relu")])
        self.fc1 = mlp([(512 + self.VEC, 512, "relu")])
        self.fc2 = nn.Linear(512, 512)

    def engine_in_ch(self):
        return self.conv1.in_channels


class NavPedPreNet(NavPreNet):
    """Same stack on cat(state[0], state[2]) (sensor map + 3-channel pedestrian map), nav_encoder.py:46-79."""
    ARCH = "navped"

    def __init__(self, image_channel=4, last_output_dim=512):
        super().__init__(image_channel, last_output_dim)


class NavPreNet1D(PreNet):
    """laser: conv1d(1->32,k5,s2) -> conv1d(32->32,k3,s2) (no activation between) -> 7616 -> fc_1d+relu (256);
    ped map: [conv k7/k5/k3 p1 + relu + maxpool2] (48->44->22->20->10->10->5) -> 6400 -> fc0+relu;
    cat(laser256, img512, vec5) -> fc1+relu -> fc2.   nav_encoder.py:82-128."""
    ARCH = "nav1d"
    VEC = 5

    def __init__(self, image_channel=1, last_output_dim=512, laser_channel=1):
        # laser_channel: 1 in the reference (nav_encoder.py:87); 3 = the non-reference "3 x 960" bench variant
        super().__init__()
        _conv_stack(self, image_channel, (7, 5, 3))
        self.laser_channel = int(laser_channel)
        self.conv1d1 = nn.Conv1d(self.laser_channel, 32, 5, 2, "valid")
        self.conv1d2 = nn.Conv1d(32, 32, 3, 2, "valid")
        self.fc_1d = mlp([(32 * 238, 256, "relu")])
        self.fc0 = mlp([(256 * 5 * 5, 512, "relu")])
        self.fc1 = mlp([(256 + 512 + self.VEC, 512, "relu")])
        self.fc2 = nn.Linear(512, 512)

    def engine_in_ch(self):
        return self.conv1.in_channels
