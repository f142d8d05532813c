"""Consumers of the forward kernels outside SHAC: policy evaluation and the no-grad rollout collection of
the model-free algorithms.

* ``evaluate_policy``  -- reference ``Logger.evaluate_policy`` (src/logger.py:91-131) without the rendering: deterministic
  actions from ``eval_reset()`` until truncation.  The reference reads ``reward.sum().item()`` and ``terminal[0].item()`` on
  the host every step (Q13); truncation depends only on the shared step counter, so here the loop length is known on the
  host and the reward is accumulated on the device: ONE synchronisation per evaluation.
* ``RolloutCollector`` -- reference ``PPO.collect_trajectories`` / ``SAC`` interaction loops (src/learning_algorithms/
  ppo.py:99-120, sac.py:110-132): ``len_trajectories`` no-grad steps of policy + safeguarded environment + critic into a
  ``CoupledBuffer`` with log-probabilities.  When the environment / policy pair has a whole-horizon engine the collection is
  ONE forward launch of it (no activation stash: nothing is differentiated) plus one batched critic launch; otherwise the
  per-step kernels run under ``torch.no_grad``.  Interventions are counted on the device (safeguard.py:76-80).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .components.coupled_buffer import CoupledBuffer


def evaluate_policy(env, policy, max_steps: int | None = None) -> float:
    """Average reward per environment and step of the deterministic policy on the evaluation episode."""
    with torch.no_grad():
        observation, _ = env.eval_reset()
        base = env.env if hasattr(env, "consts") else env
        steps = int(base.num_steps) if max_steps is None else int(max_steps)      # truncation: steps >= num_steps for every env
        total = torch.zeros((), dtype=observation.dtype, device=observation.device)
        for _ in range(steps):
            action = policy.predict(observation, deterministic=True)
            observation, reward, terminated, truncated, _ = env.step(action)
            total = total + reward.sum()
        return float(total) / base.num_envs / max(steps, 1)


class RolloutCollector:
    """No-grad rollout collection for model-free consumers.  ``policy`` / ``value_function`` are the repo's modules; the
    object quacks like the part of a learning algorithm the whole-horizon engine needs (env, policy, critic, buffer)."""

    def __init__(self, env, policy, value_function, len_trajectories: int, fused: str = "auto", rng_order: str = "reference"):
        self.env, self.policy, self.len_trajectories = env, policy, len_trajectories
        self.value_function = self.target_value_function = value_function
        self.buffer = CoupledBuffer(len_trajectories + 1, env.num_envs, env.obs_dim, True, env.action_dim, store_log_probs=True)
        self.policy_optim = None                       # nothing is updated here
        self.GAMMA = 0.99
        self.fused, self.rng_order = fused, rng_order
        self._fused_engine = None
        observation, _ = env.reset()
        with torch.no_grad():
            self.buffer.reset(observation, value_function(observation).squeeze(dim=1))

    def _engine(self):
        if self.fused == "off":
            return None
        if self._fused_engine is None:
            from .fused_rollout import FusedRollout
            eng = FusedRollout.try_create(self)
            if eng is None:
                if self.fused == "on":
                    raise RuntimeError("fused='on' but this env/safeguard/network combination has no whole-horizon engine")
                self.fused = "off"
                return None
            eng.rng_mode, eng.forward_only, eng.use_graphs = self.rng_order, True, False
            self._fused_engine = eng
        return self._fused_engine

    def collect(self) -> float:
        """One horizon into the buffer; returns the average reward (ppo.py:99-120)."""
        self.buffer.reset()
        eng = self._engine()
        H = self.len_trajectories
        with torch.no_grad():
            if eng is not None:
                avg = eng.collect_trajectories()
                obs, act = self.buffer.observations.tensor, self.buffer.actions.tensor
                flat_obs = obs[:H].reshape(-1, obs.shape[2])
                lp = self.policy.log_prob(act[:H].reshape(-1, act.shape[2]), flat_obs).view(H, -1)
                self.buffer.log_probs.tensor = torch.cat([lp, lp.new_zeros(obs.shape[0] - H, lp.shape[1])], dim=0)
                return avg
            total = torch.zeros(())
            for t in range(H):
                observation = self.buffer.observations[self.buffer.t]
                action = self.policy(observation)
                log_prob = self.policy.log_prob(action, observation)
                nxt, reward, terminated, truncated, _ = self.env.step(action)
                value = self.value_function(nxt).squeeze(dim=1)
                terminal = terminated | truncated
                if t == H - 1:
                    terminal = torch.ones_like(terminal)
                self.buffer.add(nxt, reward, terminal, value, action, log_prob)
                total = total + reward.sum()
            return float(total) / self.env.num_envs / H
