"""NavigateSeeker (reference: src/envs/navigate_seeker.py): reach a goal around circular obstacles.

Kernel side (csrc): dynamics + elastic collision operator + clamp + reward (sgbpo_step.cuh), the per-step safe
action set P5 (sgbpo_safeguard2.cuh::seeker_action_set).  Host side (here): the rejection-sampling resets
(navigate_seeker.py:107-176, RNG draws in the reference's order) and the constant obstacle columns of the
observation.  Rendering is out of scope."""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from .. import _lib, ops, rng, sets
from ..consts import build_consts
from .interfaces.obstacle_course import ObstacleCourse
from .interfaces.safe_action_env import SafeActionEnv
from .simulators.seeker import SeekerEnv

MAXK = 3      # obstacles the kernels carry per environment (csrc/sgbpo_env.cuh: SEEKER_MAXK)


class NavigateSeekerEnv(ObstacleCourse, SeekerEnv, SafeActionEnv):
    ENV_NAME = "NavigateSeeker"
    EVAL_ENVS: int = 6
    DISTANCE_PENALTY: float = 2.0
    DYNAMIC_ACTION_SET = True

    def __init__(self, num_envs: int, num_steps: int, num_obstacles: int, min_radius: float, max_radius: float,
                 draw_safe_action_set: bool = False):
        if num_obstacles > MAXK:
            raise NotImplementedError(f"the kernels carry at most {MAXK} obstacles per environment")
        SafeActionEnv.__init__(self, num_action_gens=2)
        self._init_obstacles(num_envs, num_obstacles, min_radius, max_radius)
        SeekerEnv.__init__(self, num_envs, num_steps)
        self.num_obstacles, self.min_radius, self.max_radius = num_obstacles, min_radius, max_radius
        self.draw_safe_action_set = draw_safe_action_set
        self.obs_dim = 4 + 3 * num_obstacles
        self.obstacles = [sets.Ball(torch.zeros((num_envs, 2)), torch.ones(num_envs)) for _ in range(num_obstacles)]
        self.collided = torch.zeros((num_envs, num_obstacles), dtype=torch.bool)
        self.reached = torch.zeros(num_envs, dtype=torch.bool)
        self.initial_goal_distance = torch.zeros(num_envs)
        for k in (self._plain, self._raw):
            k.n_obstacles = num_obstacles
        self.last_safe_action_set = sets.Zonotope(torch.zeros(num_envs, 2), torch.zeros(num_envs, 2, 2))

    # ---- kernel inputs ---------------------------------------------------------------------------------
    def _aux(self) -> Tensor:
        """(N, 2 + 3*MAXK): goal, then (cx, cy, r) per obstacle (unused slots zero)."""
        cols = [self.goal]
        for b in self.obstacles:
            cols += [b.center, b.radius.unsqueeze(1)]
        pad = 2 + 3 * MAXK - sum(c.shape[1] for c in cols)
        if pad:
            cols.append(torch.zeros(self.num_envs, pad, dtype=self.goal.dtype))
        return torch.cat([c.to(self.state.dtype) for c in cols], dim=1).detach()

    def _wrap_obs(self, obs4: Tensor) -> Tensor:
        return torch.cat([obs4, self._obstacle_columns().to(obs4.dtype)], dim=1)

    # ---- reference API -------------------------------------------------------------------------------------
    def _after_reset(self):
        self.goal = self.state[:, 0:2].detach().clone()
        self.sample_goal()
        self.sample_obstacles()
        self.collided = torch.zeros_like(self.collided)
        self.reached = torch.zeros_like(self.reached)

    def reset(self, seed=None):
        obs4, info = super().reset(seed)
        return obs4, info

    @property
    def observation(self) -> Tensor:
        return self._wrap_obs(SeekerEnv.observation.fget(self))

    def _step_impl(self, action: Tensor, consts, dirs):
        obs4, reward, terminated, truncated, info = super()._step_impl(action, consts, dirs)
        hit = (self.last_flags & _lib.FL_COLLIDED) != 0
        self.collided = hit.unsqueeze(1).expand(-1, self.num_obstacles) if self.num_obstacles else self.collided
        return self._wrap_obs(obs4), reward, terminated, truncated, info

    # ---- safety ---------------------------------------------------------------------------------------------
    def _set_consts(self):
        k = build_consts(self.ENV_NAME, _lib.SG_NONE, dynamic_action_set=True, n_obstacles=self.num_obstacles)
        return k

    def safe_action_set(self) -> sets.Zonotope:
        """navigate_seeker.py:398-413 (no_grad): Z(action_set.center, U diag(len)) from program P5."""
        with torch.no_grad():
            s = self.state.detach().contiguous()
            aux = self._aux().contiguous()
            gen = torch.empty(self.num_envs, 2, 2, dtype=s.dtype, device=s.device)
            status = torch.empty(self.num_envs, dtype=torch.int32, device=s.device)
            _lib.check(_lib.lib().sgbpo_seeker_action_set(_lib.dtype_code(s.dtype), self.num_envs, C.byref(self._set_consts()),
                                                         _lib.ptr(s), _lib.ptr(aux), _lib.ptr(gen), _lib.ptr(status),
                                                         _lib.stream_ptr()), "sgbpo_seeker_action_set")
            self.last_safe_action_set = sets.Zonotope(self.action_set.center.to(s.dtype), gen)
            return self.last_safe_action_set
