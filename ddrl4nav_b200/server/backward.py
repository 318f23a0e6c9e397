"""Backward module compute body -- USTC_lab/server/backward.py:182-209.

``BackwardModule.train_on(exp)`` = ``train_data.to_tensor(...)`` + ``for ... in net.learn(train_data)``;
the Redis / logger / checkpoint lines stay with the reference thread."""
from typing import Dict, List

import torch


class BackwardModule:
    def __init__(self, net, device="cuda", tensortype=torch.float32):
        self.net = net
        self.device = device
        self.tensortype = tensortype
        self.data_len = 0

    def train_on(self, train_data) -> List[Dict[str, float]]:
        train_data.to_tensor(dtype=self.tensortype, device=self.device)
        self.data_len += len(train_data)
        logs = []
        for loss_items, update_time, last in self.net.learn(train_data):
            logs.append(dict(loss_items, update_time=update_time))
        return logs
