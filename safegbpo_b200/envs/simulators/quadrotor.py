"""Planar quadrotor simulator (reference: src/envs/simulators/quadrotor.py)."""
import math

import torch

from ... import sets
from .interfaces.simulator import Simulator


class QuadrotorEnv(Simulator):
    ENV_NAME = "BalanceQuadrotor"
    DT, GRAVITY, GAIN0, GAIN1, GAIN2 = 0.05, 9.81, 70.0, 17.0, 55.0
    THRUST_MAG, ROLL_ANGLE_MAG = 6 / 7, math.pi / 12

    def __init__(self, num_envs: int = 1, num_steps: int = 1000, additional_observation_set=None):
        rep = lambda v: torch.diag_embed(torch.tensor(v).repeat(num_envs, 1))
        half = [8.0, 8.0, math.pi / 12, 0.8, 1.0, math.pi / 2]
        super().__init__(2, sets.AxisAlignedBox(torch.zeros(num_envs, 6), rep(half)),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 6), rep([0.0, 0.0, 0.0, 0.1, 0.1, 0.0])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 8), rep(half + [8.0, 8.0])), num_envs)
        self.num_steps = num_steps
        self.goal = torch.zeros((num_envs, 2))

    def _aux(self):
        return self.goal

    def _after_reset(self):
        self.goal = self.state[:, 0:2].detach().clone()

    def _adopt_reset(self, state):
        self.goal = state[:, 0:2].detach().clone()
