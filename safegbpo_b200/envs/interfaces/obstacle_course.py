"""Goal and obstacle sampling shared by the Navigate* tasks (reference: src/envs/navigate_seeker.py:107-176,
src/envs/navigate_quadrotor.py:132-196): rejection sampling on the host, RNG draws in the reference's order.

Q10: the quadrotor variant of ``sample_one_obstacle`` (navigate_quadrotor.py:177-178) assigns un-masked tensors to masked
rows and raises a shape error as soon as only SOME environments re-sample; the seeker variant (navigate_seeker.py:159-160)
masks both sides.  Both tasks use the seeker's semantics here."""
from __future__ import annotations

import torch
from torch import Tensor

from ... import rng, sets


class ObstacleCourse:
    """Mixin: expects ``state`` (positions in columns 0:2), ``goal``, ``num_envs``, ``num_obstacles``, ``min_radius``,
    ``max_radius``, ``obstacles`` (list of sets.Ball) and ``additional_observation_set``."""

    def _init_obstacles(self, num_envs: int, num_obstacles: int, min_radius: float, max_radius: float):
        mid = min_radius + (max_radius - min_radius) / 2
        self.additional_observation_set = sets.AxisAlignedBox(
            torch.tensor([0.0, 0.0, mid] * num_obstacles).unsqueeze(0).repeat(num_envs, 1),
            torch.diag_embed(torch.tensor([8.0, 8.0, (max_radius - min_radius) / 2] * num_obstacles).repeat(num_envs, 1)))

    def _reset_distribution(self):
        return None          # goal and obstacles are re-sampled by host rejection loops at every reset

    def _obstacle_columns(self) -> Tensor:
        return torch.cat([torch.cat([b.center, b.radius.unsqueeze(1)], dim=1) for b in self.obstacles], dim=1) \
            if self.obstacles else torch.zeros(self.num_envs, 0)

    def sample_goal(self):
        """navigate_seeker.py:107-115 / navigate_quadrotor.py:132-141."""
        too_close = torch.ones(self.num_envs, dtype=torch.bool)
        while bool(too_close.any()):
            self.goal[too_close] = rng.rand(int(too_close.sum()), 2).to(self.goal.dtype) * 16.0 - 8.0
            too_close = torch.linalg.vector_norm(self.goal - self.state[:, :2].detach(), dim=1) < self.max_radius * 2

    def sample_obstacles(self):
        """navigate_seeker.py:117-176 / navigate_quadrotor.py:143-196."""
        for i in range(self.num_obstacles):
            obstructing = torch.ones(self.num_envs, dtype=torch.bool)
            while bool(obstructing.any()):
                self._sample_one_obstacle(i, obstructing)
                obstructing = self._check_obstruction(i)

    def _sample_one_obstacle(self, i: int, obstructing: Tensor):
        sample = self.additional_observation_set.sample()
        pos = self.state[:, :2].detach()
        if i == 0:
            diff = self.goal - pos
            ray = diff / torch.linalg.vector_norm(diff, dim=1, keepdim=True)
            normal_ray = torch.stack([ray[:, 1], -ray[:, 0]], dim=1)
            center = (pos + self.goal) / 2 + rng.rand(self.num_envs, 1) * normal_ray * self.min_radius / 2
            radius = rng.rand(self.num_envs) * (self.max_radius - self.min_radius) + self.min_radius
        else:
            center, radius = sample[:, i * 3:i * 3 + 2], sample[:, i * 3 + 2]
        ball = self.obstacles[i]
        ball.center = torch.where(obstructing.unsqueeze(1), center.to(ball.center.dtype), ball.center)
        ball.radius = torch.where(obstructing, radius.to(ball.radius.dtype), ball.radius)

    def _check_obstruction(self, i: int) -> Tensor:
        ball = self.obstacles[i]
        obstructing = ball.contains(self.state[:, :2].detach()) | ball.contains(self.goal)
        for other in self.obstacles[:i]:
            obstructing |= ball.intersects(other)
        return obstructing
