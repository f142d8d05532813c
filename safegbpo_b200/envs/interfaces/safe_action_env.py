"""Marker ABC (reference: src/envs/interfaces/safe_action_env.py)."""
from abc import ABC, abstractmethod


class SafeActionEnv(ABC):
    def __init__(self, num_action_gens: int):
        self.num_action_gens = num_action_gens

    @abstractmethod
    def safe_action_set(self): ...
