"""NavigateQuadrotor (reference: src/envs/navigate_quadrotor.py): fly a planar quadrotor to a goal around circular obstacles.

Kernel side (csrc): quadrotor dynamics + elastic collision operator (position AND velocity reflected, roll kept,
navigate_quadrotor.py:275-328) + clamp + reward with the one-off goal bonus (sgbpo_step.cuh), and the per-step safe STATE set
(programs P6-P8, navigate_quadrotor.py:436-779; csrc/sgbpo_navquad.cuh), which the safeguards consume as per-environment
containment rows (state model 4).  Host side (here): the rejection-sampling resets (navigate_quadrotor.py:121-196; Q10: the
seeker's masked assignment, see interfaces/obstacle_course.py), the constant obstacle columns of the observation and the
stateful ``reached`` flag of the reward (the kernel reports a first arrival through FL_REACHED).  Rendering is out of scope."""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from .. import _lib, sets
from ..consts import build_consts
from .interfaces.obstacle_course import ObstacleCourse
from .interfaces.safe_state_env import SafeStateEnv
from .simulators.quadrotor import QuadrotorEnv

MAXK = 3      # obstacles the kernels carry per environment (csrc/sgbpo_env.cuh: SEEKER_MAXK)
NAUX = 2 + 3 * MAXK + 2


class NavigateQuadrotorEnv(ObstacleCourse, QuadrotorEnv, SafeStateEnv):
    ENV_NAME = "NavigateQuadrotor"
    EVAL_ENVS: int = 6
    COLLISION_PENALTY: float = 10.0
    DISTANCE_PENALTY: float = 1.0
    GOAL_REWARD: float = 100.0
    DYNAMIC_STATE_SET = True

    def __init__(self, num_envs: int, num_steps: int, num_obstacles: int, min_radius: float, max_radius: float,
                 draw_safe_state_set: bool = False):
        if num_obstacles > MAXK:
            raise NotImplementedError(f"the kernels carry at most {MAXK} obstacles per environment")
        SafeStateEnv.__init__(self, num_state_gens=6)
        self._init_obstacles(num_envs, num_obstacles, min_radius, max_radius)
        QuadrotorEnv.__init__(self, num_envs, num_steps)
        self.num_obstacles, self.min_radius, self.max_radius = num_obstacles, min_radius, max_radius
        self.draw_safe_state_set = draw_safe_state_set
        self.obs_dim = 8 + 3 * num_obstacles
        self.obstacles = [sets.Ball(torch.zeros((num_envs, 2)), torch.ones(num_envs)) for _ in range(num_obstacles)]
        self.collided = torch.zeros((num_envs, num_obstacles), dtype=torch.bool)
        self.initial_goal_distance = torch.ones(num_envs)
        self.reached = torch.zeros(num_envs, dtype=torch.bool)
        for k in (self._plain, self._raw):
            k.n_obstacles = num_obstacles
        self.last_safe_state_set = sets.Zonotope(torch.zeros(num_envs, 6), torch.zeros(num_envs, 6, 6))
        self._set_k = None

    # ---- kernel inputs ---------------------------------------------------------------------------------
    def _aux(self) -> Tensor:
        """(N, 13): goal, (cx, cz, r) per obstacle (unused slots zero), initial goal distance, reached flag."""
        dt = self.state.dtype
        cols = [self.goal.to(dt)]
        for b in self.obstacles:
            cols += [b.center.to(dt), b.radius.to(dt).unsqueeze(1)]
        pad = 2 + 3 * MAXK - sum(c.shape[1] for c in cols)
        if pad:
            cols.append(torch.zeros(self.num_envs, pad, dtype=dt))
        cols += [self.initial_goal_distance.to(dt).unsqueeze(1), self.reached.to(dt).unsqueeze(1)]
        return torch.cat(cols, dim=1).detach()

    def _wrap_obs(self, obs8: Tensor) -> Tensor:
        return torch.cat([obs8, self._obstacle_columns().to(obs8.dtype)], dim=1)

    # ---- reference API -------------------------------------------------------------------------------------
    def _after_reset(self):
        """navigate_quadrotor.py:121-129 after Simulator.reset / QuadrotorEnv.reset (goal = position)."""
        self.goal = self.state[:, 0:2].detach().clone()
        self.sample_goal()
        self.sample_obstacles()
        self.collided = torch.zeros_like(self.collided)
        self.reached = torch.zeros_like(self.reached)
        self.initial_goal_distance = torch.linalg.vector_norm(self.goal - self.state[:, :2].detach(), dim=1)

    @property
    def observation(self) -> Tensor:
        return self._wrap_obs(QuadrotorEnv.observation.fget(self))

    def _step_impl(self, action: Tensor, consts, dirs):
        obs8, reward, terminated, truncated, info = super()._step_impl(action, consts, dirs)
        flags = self.last_flags
        hit = (flags & _lib.FL_COLLIDED) != 0
        self.collided = hit.unsqueeze(1).expand(-1, self.num_obstacles) if self.num_obstacles else self.collided
        self.reached = self.reached | ((flags & _lib.FL_REACHED) != 0)             # navigate_quadrotor.py:230-231
        return self._wrap_obs(obs8), reward, terminated, truncated, info

    def eval_reset(self):
        """navigate_quadrotor.py:330-367: six fixed scenarios when num_obstacles == 1 and num_envs == 6."""
        if self.num_obstacles == 1 and self.num_envs == 6:
            self.reset()
            dt = self.state.dtype
            state = self.state.detach().clone()
            state[:, 0:2] = torch.tensor([[-4.0, 0.0], [4.0, 0.0], [0.0, 4.0], [0.0, -4.0], [-4.0, 1.0], [4.0, 1.0]], dtype=dt)
            self.state = state
            self.goal = torch.tensor([[4.0, 0.0], [-4.0, 0.0], [0.0, -4.0], [0.0, 4.0], [4.0, 1.0], [-4.0, 0.0]], dtype=dt)
            center = torch.zeros(6, 2, dtype=dt)
            center[-1, :] = 1.0
            radius = torch.full((6,), float(self.min_radius), dtype=dt)
            radius[-1] = 1.0
            self.obstacles[0].center, self.obstacles[0].radius = center, radius
            return self.observation, {}
        return super().eval_reset()

    # ---- safety ---------------------------------------------------------------------------------------------
    def _state_set_consts(self):
        if self._set_k is None:
            _, _, _, E = self.linear_dynamics()
            G_w = (E[0].double() @ self.noise_set.generator[0].double()).cpu().numpy()
            self._set_k = build_consts(self.ENV_NAME, _lib.SG_NONE, dynamic_state_set=True, n_obstacles=self.num_obstacles, G_w=G_w)
        return self._set_k

    def safe_state_set(self) -> sets.Zonotope:
        """navigate_quadrotor.py:436-460 (no_grad): Z(center, [position | velocity | roll generators]) from P6-P8.  Environments
        whose set degenerates (the reference's all-rows 1e-2 fallback blocks) get that fallback here too."""
        with torch.no_grad():
            s = self.state.detach().contiguous()
            aux = self._aux().contiguous()
            center = torch.empty(self.num_envs, 6, dtype=s.dtype, device=s.device)
            gen = torch.empty(self.num_envs, 6, 6, dtype=s.dtype, device=s.device)
            status = torch.empty(self.num_envs, dtype=torch.int32, device=s.device)
            _lib.check(_lib.lib().sgbpo_navquad_safe_state_set(_lib.dtype_code(s.dtype), self.num_envs, C.byref(self._state_set_consts()),
                                                              _lib.ptr(s), _lib.ptr(aux), _lib.ptr(center), _lib.ptr(gen),
                                                              _lib.ptr(status), _lib.stream_ptr()), "sgbpo_navquad_safe_state_set")
            self.last_safe_state_status = status
            self.last_safe_state_set = sets.Zonotope(center, gen)
            return self.last_safe_state_set
