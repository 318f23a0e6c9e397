"""MLPPreNet (classical / mujoco tasks) -- USTC_lab/nn/mlp_encoder.py:12-29: relu(linear(state[0]))."""
from .base import PreNet
from .utils import mlp


class MLPPreNet(PreNet):
    ARCH = "mlp"

    def __init__(self, input_dim=4, last_output_dim=128):
        super().__init__()
        self.fc0 = mlp([(input_dim, last_output_dim, "relu")])

    def engine_in_ch(self):
        return self.fc0[0].in_features
