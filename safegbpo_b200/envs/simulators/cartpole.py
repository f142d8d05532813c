"""CartPole simulator (reference: src/envs/simulators/cartpole.py)."""
import math

import torch

from ... import sets
from .interfaces.simulator import Simulator


class CartPoleEnv(Simulator):
    ENV_NAME = "BalanceCartPole"
    DT, GRAVITY, HALF_LENGTH, MASS_CART, MASS_POLE, FORCE_MAG = 0.02, 9.8, 0.5, 1.0, 0.1, 10.0

    def __init__(self, num_envs: int, num_steps: int):
        rep = lambda v: torch.diag_embed(torch.tensor(v).repeat(num_envs, 1))
        super().__init__(1, sets.AxisAlignedBox(torch.zeros(num_envs, 4), rep([10.0, math.pi, 10.0, 10.0])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 4), rep([0.0, 0.0, 0.0, 0.1])),
                         sets.AxisAlignedBox(torch.zeros(num_envs, 5), rep([10.0, 1.0, 1.0, 10.0, 10.0])), num_envs)
        self.num_steps = num_steps
