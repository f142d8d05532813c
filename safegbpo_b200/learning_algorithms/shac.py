"""Short Horizon Actor Critic (reference: src/learning_algorithms/shac.py:14-251).

``collect_trajectories`` + the backward half of ``update_policy`` are the hot path.  Two execution modes
with identical results:
  * eager  -- the reference's python loop over the horizon; each step is the policy MLP (torch) plus ONE
              fused CUDA launch for safeguard + simulator (ops.SafeguardedStepFn);
  * fused  -- the whole horizon in csrc/sgbpo_rollout.cu (see ``fused_rollout.py``), used automatically
              when env, safeguard and networks are of the supported fused types.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from .components.coupled_buffer import CoupledBuffer
from .components.value_function import ValueFunction
from .interfaces.learning_algorithm import LearningAlgorithm


class SHAC(LearningAlgorithm):
    GAMMA = 0.99
    MAX_GRAD_NORM = 1.0

    def __init__(self, env, policy_kwargs: dict, policy_optim_kwargs: dict, vf_kwargs: dict, vf_optim_kwargs: dict,
                 regularisation_coefficient: float, len_trajectories: int, polyak_target: float = 0.995,
                 td_weight: float = 0.95, vf_num_fits: int = 16, vf_fit_num_batches: int = 4, fused: str = "auto",
                 rng_order: str = "reference"):
        super().__init__(env, policy_kwargs, policy_optim_kwargs, vf_kwargs, vf_optim_kwargs,
                         regularisation_coefficient, False)
        self.len_trajectories = len_trajectories
        self.polyak_target, self.td_weight = polyak_target, td_weight
        self.vf_fit_num_batches, self.vf_num_fits = vf_fit_num_batches, vf_num_fits
        self.target_value_function = ValueFunction(self.env.obs_dim, **vf_kwargs)
        # Q3: the reference probes `safe_actions`, an attribute that never exists -> regularisation is dead code
        self.buffer = CoupledBuffer(len_trajectories + 1, self.env.num_envs, self.env.obs_dim, True,
                                    self.env.action_dim, store_safe_actions=hasattr(self.env, "safe_actions"))
        reset_observation, _ = self.env.reset()
        reset_value = self.target_value_function(reset_observation).squeeze(dim=1)
        self.buffer.reset(reset_observation, reset_value)
        self.interactions_per_episode = len_trajectories * self.env.num_envs
        self.fused = fused
        self.rng_order = rng_order      # "reference": draws in the reference's call order; "batched": 3 draws per episode
        self._fused_engine = None
        # data-parallel hook: called between backward and clip/Adam (distributed.allreduce_policy_grads)
        self.grad_sync = None
        self._fused_tail = None          # resolved on first use: fused clip + Adam launch when the optimizer is plain Adam on CUDA

    # ---- episode ---------------------------------------------------------------------------------------
    def _learn_episode(self, eps: int) -> tuple[float, float, float]:
        self.buffer.reset()
        self.policy_optim.zero_grad()
        average_reward = self.collect_trajectories()
        policy_loss = self.update_policy()
        value_loss = self.update_value_function()
        self.update_target_value_function()
        return average_reward, policy_loss, value_loss

    def _engine(self):
        if self.fused == "off":
            return None
        if self._fused_engine is None:
            from .fused_rollout import FusedRollout
            self._fused_engine = FusedRollout.try_create(self)
            if self._fused_engine is not None:
                self._fused_engine.rng_mode = self.rng_order
            if self._fused_engine is None and self.fused == "on":
                raise RuntimeError("fused='on' but this env/safeguard/network combination has no fused rollout")
            if self._fused_engine is None:
                self.fused = "off"
        return self._fused_engine

    def collect_trajectories(self) -> float:
        engine = self._engine()
        if engine is not None:
            return engine.collect_trajectories()
        self.env.cut_computation_graph()
        reward_sum = torch.zeros(())
        for t in range(self.len_trajectories):
            action = self.policy(self.buffer.observations[self.buffer.t])
            observation, reward, terminated, truncated, info = self.env.step(action)
            terminal = terminated | truncated
            value = self.target_value_function(observation).squeeze(dim=1)
            if t == self.len_trajectories - 1:
                terminal = torch.ones_like(terminal)
            self.buffer.add(observation, reward, terminal, value, action, safe_action=None)
            reward_sum = reward_sum + reward.detach().sum()        # one host sync per episode, not per step (Q13)
        return float(reward_sum) / self.env.num_envs / self.len_trajectories

    def discount_weights(self) -> tuple[Tensor, Tensor]:
        """shac.py:131-139: per-row discount with restarts after terminals, and the normaliser."""
        terminals = self.buffer.terminals.tensor
        rows = self.len_trajectories + 2
        exponent = torch.arange(rows).view(-1, 1).repeat(1, self.env.num_envs)
        # shac.py:133-134 subtracts (end + 1) from every later row for EACH terminal row `end`, cumulatively
        # (so with two or more terminals before a row the exponent goes below the restart value -- kept as
        # shipped); an exclusive cumulative sum does the same without the python loop over nonzero()
        drops = torch.cumsum(torch.where(terminals, exponent + 1, torch.zeros_like(exponent)), dim=0)
        exponent = exponent - torch.cat([torch.zeros_like(drops[:1]), drops[:-1]], dim=0)
        discount = self.GAMMA ** exponent.to(torch.get_default_dtype())
        normalisation = self.env.num_envs * self.len_trajectories + terminals.count_nonzero()
        return discount, normalisation

    def policy_loss(self) -> Tensor:
        discount, normalisation = self.discount_weights()
        values = torch.where(self.buffer.terminals.tensor, self.buffer.values.tensor, self.buffer.rewards.tensor)
        loss = -(discount * values).sum() / normalisation
        if self.buffer.store_safe_actions:
            loss = loss + self.regularisation_coefficient * torch.nn.functional.mse_loss(
                self.buffer.safe_actions.tensor, self.buffer.actions.tensor)
        return loss

    def update_policy(self) -> float:
        engine = self._engine()
        if engine is not None:
            loss = engine.backward()
        else:
            loss = self.policy_loss()
            loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self)
        if self._fused_tail is None:
            from .. import optim as sgoptim
            self._fused_tail = sgoptim.supported(self.policy_optim)
        if self._fused_tail:       # shac.py:148-149 as one launch (csrc/sgbpo_optim.cu)
            from .. import optim as sgoptim
            sgoptim.clip_adam_step(self.policy_optim, self.MAX_GRAD_NORM)
        else:
            torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.MAX_GRAD_NORM)
            self.policy_optim.step()
        return float(loss)

    # ---- critic (next after the hot path; plain torch, vectorised over environments) -------
    def calculate_estimated_values(self) -> tuple[Tensor, Tensor]:
        """TD-lambda targets, shac.py:187-241.  All environments share the time index, so the reference's
        masked scatter loop becomes a backward scan over rows; output ORDER follows the reference
        (row H-1 first, environments ascending, non-terminal rows only)."""
        H, N = self.len_trajectories, self.env.num_envs
        rewards, values, terminals = self.buffer.rewards.tensor, self.buffer.values.tensor, self.buffer.terminals.tensor
        td = torch.ones(N)
        avg = torch.zeros(N)
        term_ret = torch.zeros(N)
        est, obs = [], []
        for t in range(H - 1, -1, -1):
            estimate = ~terminals[t]
            init = terminals[t + 1] & estimate
            update = (~terminals[t + 1]) & estimate
            td = torch.where(init, torch.ones_like(td), td)
            avg = torch.where(init, torch.zeros_like(avg), avg)
            term_ret = torch.where(init, rewards[t] + self.GAMMA * values[t + 1], term_ret)
            td_u = td * self.td_weight
            avg_u = avg * self.td_weight * self.GAMMA + self.GAMMA * values[t + 1] \
                + (1 - td_u) / (1 - self.td_weight) * rewards[t]
            td = torch.where(update, td_u, td)
            avg = torch.where(update, avg_u, avg)
            term_ret = torch.where(update, rewards[t] + self.GAMMA * term_ret, term_ret)
            e = (1 - self.td_weight) * avg + td * term_ret
            est.append(e[estimate])
            obs.append(self.buffer.observations.tensor[t][estimate])
        est, obs = torch.cat(est), torch.cat(obs)
        total = H * N
        pad = total - est.shape[0]
        if pad:
            est = torch.cat([est, torch.zeros(pad)])
            obs = torch.cat([obs, torch.zeros(pad, self.env.obs_dim)])
        return est, obs

    def update_value_function(self) -> float:
        with torch.no_grad():
            estimated_values, observations = self.calculate_estimated_values()
        batch_size = math.ceil(len(estimated_values) / self.vf_fit_num_batches)
        value_loss = torch.zeros(())
        for _ in range(self.vf_num_fits):
            for b in range(self.vf_fit_num_batches):
                lo, hi = b * batch_size, min((b + 1) * batch_size, len(estimated_values))
                if lo >= len(estimated_values):
                    break
                self.value_function_optim.zero_grad()
                pred = self.value_function(observations[lo:hi]).squeeze(dim=1)
                value_loss = (pred - estimated_values[lo:hi]).square().mean()
                value_loss.backward()
                torch.nn.utils.clip_grad_norm_(self.value_function.parameters(), self.MAX_GRAD_NORM)
                self.value_function_optim.step()
        return float(value_loss)

    def update_target_value_function(self):
        with torch.no_grad():
            for target, param in zip(self.target_value_function.parameters(), self.value_function.parameters()):
                target.mul_(self.polyak_target).add_((1 - self.polyak_target) * param)
