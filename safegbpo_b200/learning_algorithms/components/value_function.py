"""Critic (reference: src/learning_algorithms/components/value_function.py:9-66)."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .policy import build_mlp


class ValueFunction(nn.Module):
    def __init__(self, input_dim: int, net_arch: Optional[list] = None, activation_fn: str = "nn.ELU()",
                 layer_norm: bool = True):
        super().__init__()
        self.net_arch = list(net_arch) if net_arch is not None else [64, 64]
        self.activation_fn, self.layer_norm = activation_fn, layer_norm
        self.model = build_mlp([input_dim] + self.net_arch, 1, activation_fn, layer_norm)

    def forward(self, x: Tensor, a: Optional[Tensor] = None) -> Tensor:
        if a is not None:
            x = torch.cat([x, a], dim=1)
        return self.model(x)
