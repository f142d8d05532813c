"""Household energy simulator (reference: src/envs/simulators/household.py)."""
from pathlib import Path

import numpy as np
import torch

from ... import _lib, sets
from .interfaces.simulator import Simulator

_ASSETS = Path(__file__).resolve().parent.parent.parent / "assets" / "assets.npz"


class HouseholdEnv(Simulator):
    ENV_NAME = "ManageHousehold"
    START_TIME = 7800

    def __init__(self, num_envs: int, num_steps: int):
        rep = lambda v: torch.diag_embed(torch.tensor(v).repeat(num_envs, 1))
        ctr = lambda v: torch.tensor(v).unsqueeze(0).repeat(num_envs, 1)
        super().__init__(2, sets.AxisAlignedBox(ctr([5.0, 21.0, 55.0]), rep([5.0, 3.0, 45.0])),
                         sets.AxisAlignedBox(ctr([0.0, 15.0, 0.0]), rep([0.0, 25.0, 0.0])),
                         sets.AxisAlignedBox(ctr([5.0, 21.0, 55.0] + [0.5] * 10 + [15.0] * 5 + [0.5] * 5),
                                             rep([5.0, 3.0, 45.0] + [0.5] * 10 + [25.0] * 5 + [0.5] * 5)), num_envs)
        self.num_steps = num_steps
        data = np.load(_ASSETS)
        s0 = self.START_TIME
        self._load = data["load_data"][s0:]
        self._pv = data["pv_data"][s0:]
        self._t_out = data["outside_temperature_data"][s0:]
        self._price = np.tile(np.array([0.2] * 6 + [0.4] * 5 + [0.2] * 6 + [0.5] * 5 + [0.2] * 2), 366)[s0:]

    def _shared(self):
        t = self.steps
        sh = _lib.StepShared()
        vals = np.concatenate([self._load[t:t + 5], self._pv[t:t + 5], self._t_out[t:t + 5], self._price[t:t + 5]])
        for i, v in enumerate(vals):
            sh.series[i] = float(v)
        sh.load_t, sh.pv_t, sh.price_t = float(self._load[t]), float(self._pv[t]), float(self._price[t])
        return sh

    def _noise(self):
        noise = self.noise_set.sample()
        noise[:, 1] = float(self._t_out[self.steps])      # household.py:175-176
        return noise
