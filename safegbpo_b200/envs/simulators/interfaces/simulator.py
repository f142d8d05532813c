"""Vectorised simulator base with the reference's surface
(src/envs/simulators/interfaces/simulator.py:13-282) on top of the fused CUDA step.

Differences that are deliberate and documented in DESIGN.md:
* there is no ``unbatched_dynamics`` python function to vmap: the dynamics live in
  csrc/sgbpo_env.cuh; ``dynamics(state, action, noise)`` calls the same kernel with the clamp disabled;
* ``linear_dynamics()`` returns the kernel's dual-number Jacobians (forward values; gradients through
  the linearisation are produced inside the fused step's backward, not through this accessor);
* ``scheduled_reset`` is mirrored by a host bool (truncation depends only on the shared step counter,
  simulator.py:102-106, so no device->host sync is needed to decide the reset-all of Q6);
* ``state`` before the first ``reset()`` is zeros (the reference leaves torch.empty garbage, Q4).
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch
from torch import Tensor

from abc import ABC

from .... import _lib, ops, rng
from .... import sets
from ....consts import build_consts


class Simulator(ABC):             # (an ABC like the reference's: safeguard.py:229 registers Safeguard as a virtual subclass)
    EVAL_ENVS: int = 1
    ENV_NAME: str = ""            # key into _lib.ENV_IDS

    def __init__(self, action_dim: int, state_set: sets.AxisAlignedBox, noise_set: sets.AxisAlignedBox,
                 observation_set: sets.AxisAlignedBox, num_envs: int):
        self.action_dim = action_dim
        self.state_dim = state_set.center.shape[1]
        self.obs_dim = observation_set.center.shape[1]
        self.num_envs = num_envs
        self.action_set = sets.AxisAlignedBox(torch.zeros(num_envs, action_dim),
                                              torch.diag_embed(torch.ones(num_envs, action_dim)))
        self.state_set, self.noise_set, self.observation_set = state_set, noise_set, observation_set
        self.state: Tensor = torch.zeros((num_envs, self.state_dim))
        self.steps: int = 0
        self.scheduled_reset: Tensor = torch.ones(num_envs, dtype=torch.bool)
        self._reset_pending: bool = True
        self.env_id = _lib.ENV_IDS[self.ENV_NAME]
        self._plain = build_consts(self.ENV_NAME, _lib.SG_NONE)
        self._raw = build_consts(self.ENV_NAME, _lib.SG_NONE)
        self._raw.disable_clamp = 1
        self.last_flags: Optional[Tensor] = None

    # ---- hooks the task envs override -------------------------------------------------------------
    def _aux(self) -> Optional[Tensor]:
        """Per-environment constants of observation/reward that are not state (goal)."""
        return None

    def _shared(self) -> Optional[_lib.StepShared]:
        """Per-step scalars shared by all environments (household series)."""
        return None

    def _noise(self) -> Tensor:
        return self.noise_set.sample()

    def _after_reset(self):
        """Task-specific part of reset() (e.g. state = rci.sample(), goal = state[:, :2])."""

    # ---- helpers of the fused engine: reset states are drawn ahead of time (batched RNG mode) ---------------
    _RESET_ATTRS = ("state", "steps", "_reset_pending", "scheduled_reset", "goal")

    def _reset_snapshot(self) -> dict:
        return {k: getattr(self, k) for k in self._RESET_ATTRS if hasattr(self, k)}

    def _reset_restore(self, snap: dict):
        for k, v in snap.items():
            setattr(self, k, v)

    def _reset_distribution(self):
        """The distribution reset() draws the state from, as a zonotope (center (n,), generator (n, p)) of float64 numpy arrays:
        state = center + generator (2u - 1), u ~ U[0, 1]^p -- or None when reset() is more than that (host rejection sampling).
        The whole-horizon engine draws the reset states of a horizon in its one input-generation launch from it."""
        return (self.state_set.center[0].double().cpu().numpy(), self.state_set.generator[0].double().cpu().numpy())

    def _adopt_reset(self, state: Tensor):
        """Side effects reset() has besides `state` for a reset that produced `state` (e.g. goal = position)."""

    # ---- reference API --------------------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None) -> tuple[Tensor, dict[str, Any]]:
        if seed is not None:
            torch.manual_seed(seed)
        self.state = self.state_set.sample()
        self.steps = 0
        self.scheduled_reset = torch.zeros_like(self.scheduled_reset)
        self._reset_pending = False
        self._after_reset()
        return self.observation, {}

    def step(self, action: Tensor):
        return self._step_impl(action, self._plain, None)

    def _step_impl(self, action: Tensor, consts, dirs: Optional[Tensor]):
        """simulator.py:82-108 with execute_action/observation/reward fused into one kernel launch."""
        noise = self._noise()
        prev_state = self.state
        reset_state = None
        if self._reset_pending:                      # Q6: any scheduled -> reset ALL; reward on the new state
            self.steps += 1
            self.reset()                             # draws in the reference's order (after the noise sample)
            reset_state = self.state.detach()
        else:
            self.steps += 1
        s_next, a_exec, obs, reward, flags = ops.safeguarded_step(
            prev_state, action, noise, self.env_id, consts, aux=self._aux(), dirs=dirs, reset_state=reset_state,
            shared=self._shared())
        self.state = s_next
        self.last_flags = flags
        self.last_exec_action = a_exec
        terminated, truncated = self.episode_ending()
        self.scheduled_reset = terminated | truncated
        self._reset_pending = self.steps >= self.num_steps
        return obs, reward, terminated, truncated, {}

    @property
    def observation(self) -> Tensor:
        """Observation of the current state (kernel evaluated with a zero action, outputs other than obs dropped)."""
        zeros_a = torch.zeros(self.num_envs, self.action_dim, dtype=self.state.dtype, device=self.state.device)
        zeros_w = torch.zeros_like(self.state)
        out = ops.safeguarded_step(self.state.detach(), zeros_a, zeros_w, self.env_id, self._raw, aux=self._aux(),
                                   reset_state=self.state, shared=self._shared())
        return out[2]

    def reward(self, action: Tensor) -> Tensor:
        zeros_w = torch.zeros_like(self.state)
        out = ops.safeguarded_step(self.state.detach(), action, zeros_w, self.env_id, self._raw, aux=self._aux(),
                                   reset_state=self.state, shared=self._shared())
        return out[3]

    def episode_ending(self) -> tuple[Tensor, Tensor]:
        terminated = torch.zeros(self.num_envs, dtype=torch.bool)
        truncated = torch.full((self.num_envs,), self.steps >= self.num_steps, dtype=torch.bool)
        return terminated, truncated

    def execute_action(self, action: Tensor):
        out = ops.safeguarded_step(self.state, action, self._noise(), self.env_id, self._plain, aux=self._aux(),
                                   shared=self._shared())
        self.state = out[0]

    def dynamics(self, state: Tensor, action: Tensor, noise: Tensor) -> Tensor:
        aux = self._aux()
        return ops.safeguarded_step(state, action, noise, self.env_id, self._raw, aux=aux, shared=self._shared())[0]

    def linear_dynamics(self):
        return ops.linearise(self.state, self.env_id, self._plain)

    def reachable_set(self) -> sets.Zonotope:
        center, _, action_mat, _ = self.linear_dynamics()
        return sets.Zonotope(center, torch.bmm(action_mat, self.action_set.generator.to(action_mat.dtype)))

    def eval_reset(self):
        devices = list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []
        with torch.random.fork_rng(devices=devices):
            obs, info = self.reset(seed=42)
        return obs, info

    def cut_computation_graph(self):
        self.state = self.state.detach()

    def render(self):
        raise NotImplementedError("rendering is outside the hot path (SURVEY §2.1: visualisation only)")

    def close(self):
        pass
