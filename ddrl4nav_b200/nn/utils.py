"""`mlp` / `merge` helpers with the reference's signatures (USTC_lab/nn/utils.py:6-20)."""
from typing import List, Tuple

import torch
import torch.nn as nn

_ACTS = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid}


def merge(*tensors):
    return torch.cat(tensors, dim=-1)


def mlp(input_mlp: List[Tuple[int, int, str]]) -> nn.Sequential:
    """[(in, out, activation-name), ...] -> Sequential(Linear, Act, ...).  Parameter names are
    '<idx>.weight' / '<idx>.bias' exactly as in the reference, so 'fc0.0.weight' etc. line up."""
    seq = nn.Sequential()
    for fan_in, fan_out, act in input_mlp or []:
        seq.append(nn.Linear(fan_in, fan_out, bias=True))
        if act in _ACTS:
            seq.append(_ACTS[act]())
    return seq
