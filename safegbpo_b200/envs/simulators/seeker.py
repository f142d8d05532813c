"""Seeker (single integrator) simulator (reference: src/envs/simulators/seeker.py)."""
import torch

from ... import sets
from .interfaces.simulator import Simulator


class SeekerEnv(Simulator):
    ENV_NAME = "NavigateSeeker"
    DT = 0.1

    def __init__(self, num_envs: int = 1, num_steps: int = 1000, additional_observation_set=None):
        rep = lambda v: torch.diag_embed(torch.tensor(v).repeat(num_envs, 1))
        super().__init__(2, sets.AxisAlignedBox(torch.zeros(num_envs, 2), rep([8.0, 8.0])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 2), rep([0.05, 0.05])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 4), rep([8.0] * 4)), num_envs)
        self.num_steps = num_steps
        self.goal = torch.zeros((num_envs, 2))

    def _aux(self):
        return self.goal

    def _after_reset(self):
        self.goal = self.state[:, 0:2].detach().clone()
