"""RCI environments: constant control / RCI zonotopes (reference: src/envs/interfaces/rci_env.py:11-56)."""
from abc import ABC
from pathlib import Path

import numpy as np
import torch

from ... import sets
from .safe_action_env import SafeActionEnv
from .safe_state_env import SafeStateEnv

_ASSETS = Path(__file__).resolve().parent.parent.parent / "assets" / "assets.npz"


class RCIEnv(SafeActionEnv, SafeStateEnv, ABC):
    def __init__(self, num_envs: int, system: str):
        data = np.load(_ASSETS)
        t = lambda k: torch.from_numpy(data[f"{system}_{k}"]).to(torch.get_default_dtype()).to(torch.get_default_device())
        self.ctrl = sets.Zonotope(t("ctrl_center").expand(num_envs, -1), t("ctrl_generators").expand(num_envs, -1, -1))
        self.rci = sets.Zonotope(t("rci_center").expand(num_envs, -1), t("rci_generators").expand(num_envs, -1, -1))
        self._ctrl64 = (data[f"{system}_ctrl_center"], data[f"{system}_ctrl_generators"])
        self._rci64 = (data[f"{system}_rci_center"], data[f"{system}_rci_generators"])
        SafeActionEnv.__init__(self, self.ctrl.generator.shape[2])
        SafeStateEnv.__init__(self, self.rci.generator.shape[2])

    def safe_action_set(self) -> sets.Zonotope:
        return self.ctrl

    def safe_state_set(self) -> sets.Zonotope:
        return self.rci
