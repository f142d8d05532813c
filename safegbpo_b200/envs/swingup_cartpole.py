"""SwingUpCartPole (reference: src/envs/swingup_cartpole.py)."""
import math

from .. import rng, sets
from .interfaces.safe_state_env import SafeStateEnv
from .simulators.cartpole import CartPoleEnv


class SwingUpCartPoleEnv(CartPoleEnv, SafeStateEnv):
    ENV_NAME = "SwingUpCartPole"

    def __init__(self, num_envs: int, num_steps: int):
        SafeStateEnv.__init__(self, 4)
        CartPoleEnv.__init__(self, num_envs, num_steps)

    def _after_reset(self):
        self.state = rng.rand(self.num_envs, 4) * 0.1 - 0.05      # swingup_cartpole.py:62-63
        self.state[:, 1] = math.pi

    def _reset_distribution(self):
        import numpy as np
        return np.array([0.0, math.pi, 0.0, 0.0]), np.diag([0.05, 0.0, 0.05, 0.05])

    def safe_state_set(self) -> sets.Zonotope:
        return self.state_set
