"""BalanceCartPole (reference: src/envs/balance_cartpole.py)."""
from .interfaces.rci_env import RCIEnv
from .simulators.cartpole import CartPoleEnv


class BalanceCartPoleEnv(CartPoleEnv, RCIEnv):
    ENV_NAME = "BalanceCartPole"

    def __init__(self, num_envs: int, num_steps: int):
        RCIEnv.__init__(self, num_envs, "cartpole")
        CartPoleEnv.__init__(self, num_envs, num_steps)

    def _after_reset(self):
        self.state = self.rci.sample()

    def _reset_distribution(self):
        return self._rci64
