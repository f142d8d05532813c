"""Oracle SHAC rollout + policy loss (float64, CPU).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates src/learning_algorithms/components/policy.py:50-89, value_function.py:38-66,
coupled_buffer.py:94-172 and shac.py:96-151 (collect_trajectories + the loss/backward half of
update_policy; the optimiser step and the critic fit are outside the hot path).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn
from torch import Tensor

GAMMA = 0.99          # shac.py:23


def make_mlp(in_dim: int, hidden: list, out_dim: int) -> nn.Sequential:
    """Linear -> ELU -> LayerNorm per hidden layer, then Linear (policy.py:50-60, value_function.py:38-48)."""
    dims = [in_dim] + list(hidden)
    layers = []
    for i in range(len(dims) - 1):
        layers += [nn.Linear(dims[i], dims[i + 1]), nn.ELU(), nn.LayerNorm(dims[i + 1])]
    layers.append(nn.Linear(dims[-1], out_dim))
    return nn.Sequential(*layers)


class OraclePolicy(nn.Module):
    def __init__(self, obs_dim: int, act_dim: int, hidden: list):
        super().__init__()
        self.model = make_mlp(obs_dim, hidden, act_dim)
        self.log_std = nn.Parameter(-1.0 * torch.ones(act_dim))     # policy.py:63-65

    def forward(self, obs: Tensor, eps: Tensor) -> Tensor:
        """tanh(mean + exp(log_std) * eps): Normal.rsample() with the draw made explicit (policy.py:81-89)."""
        return torch.tanh(self.model(obs) + torch.exp(self.log_std) * eps)


class OracleValue(nn.Module):
    def __init__(self, obs_dim: int, hidden: list):
        super().__init__()
        self.model = make_mlp(obs_dim, hidden, 1)

    def forward(self, obs: Tensor) -> Tensor:
        return self.model(obs)


def discount_weights(terminals: Tensor, horizon: int) -> tuple[Tensor, Tensor]:
    """shac.py:131-139.  terminals: bool (H+2, N).  Returns (discount (H+2,N), normalisation scalar)."""
    n_rows, n_env = terminals.shape
    exponent = torch.arange(n_rows).view(-1, 1).repeat(1, n_env)
    for end, env in terminals.nonzero():
        exponent[end + 1:, env] -= end + 1
    discount = GAMMA ** exponent.to(torch.float64)
    norm = n_env * horizon + terminals.count_nonzero()
    return discount, norm


def rollout_loss(env_step: Callable, obs0: Tensor, policy: OraclePolicy, target_vf: OracleValue, horizon: int,
                 eps: Tensor, value0: Optional[Tensor] = None):
    """collect_trajectories (shac.py:96-120) followed by the loss of update_policy (shac.py:131-142).

    env_step(action, t) -> (obs, reward, terminated|truncated)  -- the (safeguarded) environment step.
    eps: (H, N, m) standard-normal draws consumed by the policy in order.
    Returns dict with loss and the buffer tensors (rows 0..H+1 as in CoupledBuffer(len+1)).
    """
    N = obs0.shape[0]
    rows = horizon + 2
    observations = [obs0.detach()]              # buffer.reset(): detached restart row (coupled_buffer.py:108-111)
    values = [torch.zeros(N, dtype=obs0.dtype) if value0 is None else value0.detach()]
    rewards, actions, terminals = [], [], [torch.zeros(N, dtype=torch.bool)]
    for t in range(horizon):
        action = policy(observations[t], eps[t])
        obs, reward, terminal = env_step(action, t)
        value = target_vf(obs).squeeze(1)
        if t == horizon - 1:
            terminal = torch.ones_like(terminal)
        observations.append(obs)
        values.append(value)
        rewards.append(reward)
        actions.append(action)
        terminals.append(terminal)
    pad = torch.zeros(N, dtype=obs0.dtype)
    rew = torch.stack(rewards + [pad, pad])                      # rows H, H+1 stay zero
    val = torch.stack(values + [pad])
    term = torch.stack(terminals + [torch.zeros(N, dtype=torch.bool)])
    discount, norm = discount_weights(term, horizon)
    picked = torch.where(term, val, rew)
    loss = -(discount * picked).sum() / norm
    obs_rows = torch.stack(observations + [torch.zeros_like(obs0)])   # (H+2, N, o): last row never written
    return {"loss": loss, "observations": obs_rows, "values": val, "rewards": rew,
            "terminals": term, "actions": torch.stack(actions), "discount": discount, "norm": norm}
