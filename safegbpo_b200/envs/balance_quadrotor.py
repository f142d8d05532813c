"""BalanceQuadrotor (reference: src/envs/balance_quadrotor.py)."""
from .interfaces.rci_env import RCIEnv
from .simulators.quadrotor import QuadrotorEnv


class BalanceQuadrotorEnv(QuadrotorEnv, RCIEnv):
    ENV_NAME = "BalanceQuadrotor"

    def __init__(self, num_envs: int, num_steps: int):
        RCIEnv.__init__(self, num_envs, "quadrotor")
        QuadrotorEnv.__init__(self, num_envs, num_steps)

    def _after_reset(self):
        self.state = self.rci.sample()
        self.goal = self.state[:, 0:2].detach().clone()

    def _reset_distribution(self):
        return self._rci64
