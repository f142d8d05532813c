"""BalancePendulum (reference: src/envs/balance_pendulum.py)."""
from .interfaces.rci_env import RCIEnv
from .simulators.pendulum import PendulumEnv


class BalancePendulumEnv(PendulumEnv, RCIEnv):
    ENV_NAME = "BalancePendulum"

    def __init__(self, num_envs: int, num_steps: int):
        RCIEnv.__init__(self, num_envs, "pendulum")
        PendulumEnv.__init__(self, num_envs, num_steps)

    def _after_reset(self):
        self.state = self.rci.sample()

    def _reset_distribution(self):
        return self._rci64
