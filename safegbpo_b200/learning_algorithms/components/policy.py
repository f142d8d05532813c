"""Stochastic tanh-Gaussian policy (reference: src/learning_algorithms/components/policy.py:13-158).

The module keeps the reference's parameter layout (``model`` = Linear/ELU/LayerNorm stack, ``log_std``)
so state dicts are interchangeable.  ``forward`` is the eager path; the fused rollout reads the
parameters directly (learning_algorithms/shac.py)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from ... import rng


def build_mlp(dims: list, out_dim: int, activation_fn: str, layer_norm: bool) -> nn.Sequential:
    layers = []
    for i in range(len(dims) - 1):
        layers += [nn.Linear(dims[i], dims[i + 1]), eval(activation_fn)]
        if layer_norm:
            layers.append(nn.LayerNorm(dims[i + 1]))
    layers.append(nn.Linear(dims[-1], out_dim))
    return nn.Sequential(*layers)


class Policy(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, stochastic: bool = True, net_arch: Optional[list] = None,
                 activation_fn: str = "nn.ELU()", layer_norm: bool = True):
        super().__init__()
        self.stochastic, self.input_dim, self.output_dim = stochastic, input_dim, output_dim
        self.net_arch = list(net_arch) if net_arch is not None else [64, 64]
        self.activation_fn, self.layer_norm = activation_fn, layer_norm
        self.model = build_mlp([input_dim] + self.net_arch, output_dim, activation_fn, layer_norm)
        self.log_std = 0
        if stochastic:
            self.log_std = nn.Parameter(-1.0 * torch.ones(output_dim))
        self.last_eps: Optional[Tensor] = None

    def forward(self, x: Tensor) -> Tensor:
        mean = self.model(x)
        if self.stochastic:
            eps = rng.randn(*mean.shape).to(mean.dtype)          # Normal(mean, std).rsample()
            self.last_eps = eps
            mean = mean + torch.exp(self.log_std) * eps
        return torch.tanh(mean)

    def predict(self, x: Tensor, deterministic: bool = True) -> Tensor:
        return torch.tanh(self.model(x)) if deterministic else self.forward(x)

    # ---- what the model-free consumers (PPO / SAC) ask of the policy: reference policy.py:112-158 ------------------------
    def _dist(self, x: Tensor) -> torch.distributions.Normal:
        return torch.distributions.Normal(self.model(x), torch.exp(self.log_std))

    def log_prob(self, action: Tensor, x: Tensor) -> Tensor:
        """Log-density of a tanh-squashed Gaussian action: change of variables inside (-1, 1), the Gaussian tail mass at the
        saturated values +-1 (policy.py:112-141)."""
        dist = self._dist(x)
        tiny = torch.finfo(action.dtype).eps
        clamped = action.clamp(min=-1.0 + tiny, max=1.0 - tiny)
        pre = 0.5 * (clamped.log1p() - (-clamped).log1p())                 # atanh
        inside = dist.log_prob(pre) - torch.log(1 - clamped ** 2)
        at_bound = torch.where(action > 0, torch.log(1 - dist.cdf(pre) + tiny), torch.log(dist.cdf(-pre) + tiny))
        saturated = torch.isclose(action.abs(), torch.ones_like(action))
        return torch.where(saturated, at_bound, inside).sum(dim=1)

    def entropy(self, x: Tensor) -> Tensor:
        """Entropy of the un-squashed Gaussian (the reference's stated approximation, policy.py:143-158)."""
        return self._dist(x).entropy().sum(dim=1)

