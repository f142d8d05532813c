"""Rollout buffer (reference: src/learning_algorithms/components/coupled_buffer.py:26-285): the part of
the API the SHAC hot path uses (reset / add and the tensors).  Sampling helpers for PPO/SAC
(`batch`, `unique_batches`, GAE) are outside the hot path."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .coupled_tensor import CoupledTensor


class CoupledBuffer:
    def __init__(self, len_trajectories: int, num_envs: int, obs_dim: int, store_values: bool = False,
                 action_dim: Optional[int] = None, store_log_probs: bool = False, store_safe_actions: bool = False,
                 batch_size: Optional[int] = None, gamma: Optional[float] = None, gae_lambda: Optional[float] = None):
        self.len_trajectories, self.num_envs, self.obs_dim = len_trajectories, num_envs, obs_dim
        self.store_values, self.action_dim = store_values, action_dim
        self.store_log_probs, self.store_safe_actions = store_log_probs, store_safe_actions
        self.batch_size, self.gamma, self.gae_lambda = batch_size, gamma, gae_lambda
        rows = len_trajectories + 1
        self.fast_reset = None          # set by the fused rollout engine while its static buffers back this object
        self.t = torch.zeros(num_envs, dtype=torch.int64)
        self.observations = CoupledTensor(rows, num_envs, obs_dim)
        if store_values:
            self.values = CoupledTensor(rows, num_envs)
        if action_dim is not None:
            self.actions = CoupledTensor(rows, num_envs, action_dim)
            if store_log_probs:
                self.log_probs = CoupledTensor(rows, num_envs)
            if store_safe_actions:
                self.safe_actions = CoupledTensor(rows, num_envs, action_dim)
        self.rewards = CoupledTensor(rows, num_envs)
        if gae_lambda is not None:
            self.advantages = CoupledTensor(rows, num_envs)
        self.terminals = CoupledTensor(rows, num_envs, dtype=torch.bool)

    def _all(self):
        names = ["observations", "values", "actions", "log_probs", "safe_actions", "rewards", "advantages", "terminals"]
        return [getattr(self, n) for n in names if hasattr(self, n)]

    def reset(self, reset_observation: Optional[Tensor] = None, reset_value: Optional[Tensor] = None) -> None:
        if self.fast_reset is not None and self.fast_reset(reset_observation, reset_value):
            return      # the fused engine owns the storage: row 0 is set in place, nothing re-allocated
        if reset_observation is None:
            reset_observation = self.observations[self.t].detach()
            if self.store_values:
                reset_value = self.values[self.t].detach()
        self.t = torch.zeros_like(self.t)
        for ct in self._all():
            ct.reset()
        self.observations[0] = reset_observation
        if self.store_values:
            self.values[0] = reset_value

    def add(self, observation: Tensor, reward: Tensor, terminal: Tensor, value: Optional[Tensor] = None,
            action: Optional[Tensor] = None, log_prob: Optional[Tensor] = None,
            safe_action: Optional[Tensor] = None) -> None:
        if (self.t >= self.len_trajectories).any():
            print("[BUFFER OVERFLOW] Overwriting oldest transitions")
            self.t = torch.where(self.t >= self.len_trajectories, torch.zeros_like(self.t), self.t)
        self.observations[self.t + 1] = observation
        if value is not None:
            self.values[self.t + 1] = value
        if action is not None:
            self.actions[self.t] = action
            if log_prob is not None:
                self.log_probs[self.t] = log_prob
            if safe_action is not None:
                self.safe_actions[self.t] = safe_action
        self.rewards[self.t] = reward
        self.terminals[self.t + 1] = terminal
        self.t = self.t + 1
