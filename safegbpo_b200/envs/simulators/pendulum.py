"""Pendulum simulator (reference: src/envs/simulators/pendulum.py; dynamics in csrc/sgbpo_env.cuh)."""
import math

import torch

from ... import sets
from .interfaces.simulator import Simulator


class PendulumEnv(Simulator):
    ENV_NAME = "BalancePendulum"
    DT, GRAVITY, LENGTH, MASS, TORQUE_MAG = 0.05, 9.81, 1.0, 1.0, 30.0

    def __init__(self, num_envs: int, num_steps: int):
        rep = lambda v: torch.diag_embed(torch.tensor(v).repeat(num_envs, 1))
        super().__init__(1, sets.AxisAlignedBox(torch.zeros(num_envs, 2), rep([math.pi, 8.0])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 2), rep([0.0, 0.1])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 3), rep([1.0, 1.0, 8.0])), num_envs)
        self.num_steps = num_steps
        self.last_torque = None
