"""Safeguard wrapper with the reference's surface (src/safeguards/interfaces/safeguard.py:21-229).

``actions`` / ``step`` run ONE fused kernel per call (linearise, gate, program, where, and for ``step``
also dynamics/clamp/observation/reward) instead of vmap(jacrev) x2 + a host CvxpyLayer batch.
Reference quirks are kept under ``gate="reference"`` (default): Q1 raw action where projectable,
Q4 init-time B0/G_w of environment 0, Q9 direction draw inside the safeguard.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch
from torch import Tensor

from ... import _lib, ops, rng
from ...consts import build_consts
from ...envs.interfaces.safe_action_env import SafeActionEnv
from ...envs.interfaces.safe_state_env import SafeStateEnv
from ...envs.simulators.interfaces.simulator import Simulator


class Safeguard:
    SG_KIND = _lib.SG_NONE

    def __init__(self, env, gate: str = "reference", strict: bool = True, **kwargs):
        self.env = env
        self.batch_dim, self.state_dim, self.action_dim = env.num_envs, env.state_dim, env.action_dim
        self.action_constrained = isinstance(env, SafeActionEnv)
        self.state_constrained = isinstance(env, SafeStateEnv)
        if self.state_constrained:
            self.safe_state_gens = env.num_state_gens
        self.safe_action_gens = env.num_action_gens if self.action_constrained else 2 * self.action_dim
        # safeguard.py:54: linearise once; env 0's matrices become constants of every program (Q4)
        self.constant_mat, self.state_mat, self.action_mat, self.noise_mat = env.linear_dynamics()
        self.gate, self.strict = gate, strict
        self.safe_action = None
        self._interventions = torch.zeros((), dtype=torch.int64)
        self._status = torch.zeros((), dtype=torch.int32)
        self._consts = None

    # ---- constants --------------------------------------------------------------------------------
    def _rm_flags(self) -> dict:
        return {}

    def consts(self):
        if self._consts is None:
            env = self.env
            f64 = lambda t: t.detach().to(torch.float64).cpu().numpy()
            sa = ss = None
            dynamic = bool(getattr(env, "DYNAMIC_ACTION_SET", False))
            if self.action_constrained and not dynamic:
                z = env.safe_action_set()
                sa = getattr(env, "_ctrl64", None) or (f64(z.center[0]), f64(z.generator[0]))
            dynamic_state = bool(getattr(env, "DYNAMIC_STATE_SET", False))      # NavigateQuadrotor: rebuilt per step in the kernels
            if self.state_constrained and not dynamic_state:
                z = env.safe_state_set()
                ss = getattr(env, "_rci64", None) or (f64(z.center[0]), f64(z.generator[0]))
            G_w = f64(self.noise_mat[0]) @ f64(env.noise_set.generator[0])
            self._consts = build_consts(env.ENV_NAME, self.SG_KIND, gate_mode=_lib.GATE[self.gate], safe_action=sa,
                                        safe_state=ss, G_w=G_w, B0=f64(self.action_mat[0]), dynamic_action_set=dynamic,
                                        n_obstacles=int(getattr(env, "num_obstacles", 0)), dynamic_state_set=dynamic_state,
                                        **self._rm_flags())
            # strict (the reference's behaviour): an infeasible program yields NaN and raises; non-strict: the raw action is
            # executed there, flagged FL_INFEASIBLE, so that long rollouts and the weights stay finite
            self._consts.infeasible_raw = 0 if self.strict else 1
        return self._consts

    def _directions(self) -> Optional[Tensor]:
        return None

    # ---- reference API --------------------------------------------------------------------------------
    def actions(self, action: Tensor) -> Tensor:
        env = self.env
        zeros_w = torch.zeros_like(env.state)
        out = ops.safeguarded_step(env.state, action, zeros_w, env.env_id, self.consts(), aux=env._aux(),
                                   dirs=self._directions(), shared=env._shared())
        self._account(out[1], out[4])
        return out[1]

    def safeguard(self, action: Tensor) -> Tensor:
        """The map itself, ungated (reference: abstract ``safeguard``; evaluated for all envs)."""
        env = self.env
        k = self._always_consts()
        out = ops.safeguarded_step(env.state, action, torch.zeros_like(env.state), env.env_id, k, aux=env._aux(),
                                   dirs=self._directions(), shared=env._shared())
        return out[1]

    def _always_consts(self):
        import copy
        k = copy.copy(self.consts())
        k.gate_mode = _lib.GATE["always"]
        return k

    def step(self, action: Tensor):
        """VectorActionWrapper.step = env.step(self.actions(action)), fused into one launch."""
        dirs = self._directions()                     # drawn before the noise sample (RNG order, App. B)
        out = self.env._step_impl(action, self.consts(), dirs)
        self._account(self.env.last_exec_action, self.env.last_flags)
        return out

    def _account(self, safe_action: Tensor, flags: Tensor):
        self.safe_action = safe_action
        self._interventions = self._interventions + ((flags & _lib.FL_INTERVENED) != 0).sum()
        bad = (flags & (_lib.FL_NONFINITE | _lib.FL_INFEASIBLE)) != 0
        if self.strict:                                # reference: host sync + ValueError every call (Q13)
            if bool(bad.any()):
                idx = bad.nonzero().flatten()[:8].tolist()
                if bool(((flags & _lib.FL_UNREDUCED) != 0).any()):
                    raise NotImplementedError(
                        f"envs {idx}: the safeguarded action is needed but this safe state set (non-invertible generator "
                        f"with more than 2 rows) has no GPU reduction yet; see DESIGN.md section 8")
                raise ValueError(f"Safe action are NaN / program infeasible for envs {idx} "
                                 f"(the reference raises here: safeguard.py:68-75 / solver error)")
        else:
            self._status = self._status | bad.any().to(torch.int32)

    @property
    def interventions(self) -> int:
        return int(self._interventions)

    @interventions.setter
    def interventions(self, v: int):
        self._interventions = torch.full((), int(v), dtype=torch.int64)

    def check_status(self):
        """Deferred form of the reference's per-step NaN check (one sync per call)."""
        if int(self._status):
            raise ValueError("a safeguarded step produced a non-finite action or an infeasible program")

    def reset(self, *, seed: int | None = None, options: dict | None = None) -> tuple[Tensor, dict[str, Any]]:
        return self.env.reset(seed=seed)

    def __getattr__(self, name: str):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def __dir__(self):
        return sorted(set(super().__dir__()) | set(dir(self.env)))


Simulator.register(Safeguard)      # safeguard.py:229: a wrapped environment is a Simulator for isinstance checks
