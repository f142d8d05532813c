"""Marker ABC (reference: src/envs/interfaces/safe_state_env.py)."""
from abc import ABC, abstractmethod


class SafeStateEnv(ABC):
    def __init__(self, num_state_gens: int):
        self.num_state_gens = num_state_gens

    @abstractmethod
    def safe_state_set(self): ...
