"""Actor-critic base (reference: src/learning_algorithms/interfaces/learning_algorithm.py:18-99)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from ..components.policy import Policy
from ..components.value_function import ValueFunction


class LearningAlgorithm(ABC):
    POLICY_LEARNING_RATE_SCHEDULE = "constant"
    VF_LEARNING_RATE_SCHEDULE = "constant"
    interactions_per_episode: int

    def __init__(self, env, policy_kwargs: dict, policy_optim_kwargs: dict, vf_kwargs: dict, vf_optim_kwargs: dict,
                 regularisation_coefficient: float, q_function: bool):
        self.env = env
        self.policy = Policy(env.obs_dim, env.action_dim, **policy_kwargs)
        self.value_function = ValueFunction(env.obs_dim + (env.action_dim if q_function else 0), **vf_kwargs)
        if torch.cuda.is_available() and next(self.policy.parameters()).is_cuda:
            # device-side step counter: same arithmetic, lets the update be replayed inside a CUDA graph
            policy_optim_kwargs = {"capturable": True, **policy_optim_kwargs}
            vf_optim_kwargs = {"capturable": True, **vf_optim_kwargs}
        self.policy_optim = torch.optim.Adam(self.policy.parameters(), **policy_optim_kwargs)
        self.value_function_optim = torch.optim.Adam(self.value_function.parameters(), **vf_optim_kwargs)
        self.regularisation_coefficient = regularisation_coefficient

    def learn(self, interactions: int, logger=None):
        episodes = interactions // self.interactions_per_episode
        pol_step = vf_step = 0.0
        if self.POLICY_LEARNING_RATE_SCHEDULE == "linear":
            pol_step = (1e-5 - self.policy_optim.param_groups[0]["lr"]) / episodes
        if self.VF_LEARNING_RATE_SCHEDULE == "linear":
            vf_step = (1e-5 - self.value_function_optim.param_groups[0]["lr"]) / episodes
        for eps in range(episodes):
            average_reward, policy_loss, value_loss = self._learn_episode(eps)
            for group in self.policy_optim.param_groups:
                group["lr"] += pol_step
            for group in self.value_function_optim.param_groups:
                group["lr"] += vf_step
            if logger is not None:
                with torch.no_grad():
                    logger.on_learning_episode(eps, average_reward, policy_loss, value_loss, episodes)

    @abstractmethod
    def _learn_episode(self, eps: int) -> tuple[float, float, float]: ...
