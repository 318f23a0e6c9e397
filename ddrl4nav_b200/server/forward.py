"""Forward module compute body -- USTC_lab/server/forward.py:128-170.

``ForwardModule`` owns what the reference's ``ForwardThread.run`` does between popping a Redis item
and pushing the replies: host arrays -> pinned staging -> device, ONE fused engine call
(encoders + heads + sampling + log-prob + value), device -> host.  Redis / EasyBytes stay with the
caller (the reference thread); INTEGRATION.md shows the 6-line patch."""
from typing import List, Optional, Sequence

import numpy as np
import torch


def forward_compute(net, batch_states: Sequence, play_mode: bool = False, draw: Optional[torch.Tensor] = None):
    """(actions, logps, values [V,B,1]) as fp32 device tensors -- forward.py:132-146."""
    return net.act(batch_states, draw=draw, play_mode=play_mode)


class ForwardModule:
    def __init__(self, net, play_mode: bool = False, nptype=np.float32, device="cuda"):
        self.net = net
        self.play_mode = play_mode
        self.nptype = nptype
        self.device = torch.device(device)
        self._pinned = {}

    def _stage(self, i: int, a) -> torch.Tensor:
        """np array (u8/f16/f32/f64, easybytes.py:21-26) -> fp32 device tensor through a reusable pinned buffer."""
        if torch.is_tensor(a):
            return a.to(self.device, torch.float32, non_blocking=True)
        a = np.asarray(a)
        key = (i, a.shape, a.dtype.str)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
            self._pinned = {k: v for k, v in self._pinned.items() if k[0] != i}
            self._pinned[key] = buf
        buf.numpy()[...] = a
        return buf.to(self.device, non_blocking=True).to(torch.float32)

    def step(self, batch_states: Sequence, draw: Optional[torch.Tensor] = None) -> List[np.ndarray]:
        """Returns [actions, logps, values] as numpy (the list ``encode_forward_return_data`` consumes)."""
        states = [self._stage(i, s) for i, s in enumerate(batch_states)]
        actions, logps, values = self.net.act(states, draw=draw, play_mode=self.play_mode)
        out = torch.cat([actions.reshape(-1), logps, values.reshape(-1)]).cpu().numpy().astype(self.nptype, copy=False)
        B = logps.shape[0]
        na = actions.numel()
        return [out[:na].reshape(tuple(actions.shape)), out[na:na + B], out[na + B:].reshape(1, B, 1)]
