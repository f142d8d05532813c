"""Oracle safeguards: Safeguard.actions, BoundaryProjection, RayMask restated in float64 on the CPU.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/safeguards/interfaces/safeguard.py:61-80,200-208, boundary_projection.py:22-57 and
ray_mask.py:63-341 statement by statement, *including* the shipped quirks of SURVEY.md §0.1
(Q1 gate order, Q2 radial map, Q4 init-time noise/action matrices, Q9 RNG draw inside the safeguard).
One convex program per environment per call, like the reference's CvxpyLayer batch.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import programs as P
from .dynamics import OracleEnv
from .sets import aab_intersects_zonotope

F64 = torch.float64


class OracleSafeguard:
    """safeguard.py:21-80.  ``kind``: "boundary_projection" | "ray_mask"."""

    def __init__(self, env: OracleEnv, kind: str = "boundary_projection", linear_projection: bool = True,
                 zonotopic_approximation: bool = True, passthrough: bool = False, gate: str = "reference",
                 strict: bool = True):
        self.env, self.kind = env, kind
        self.linear_projection = linear_projection
        self.zonotopic = zonotopic_approximation
        self.passthrough = passthrough
        self.gate = gate                    # "reference" (Q1 as shipped) | "intended" | "always"
        self.strict = strict                # False: keep NaN rows of infeasible programs instead of raising
        self._action_set = None             # per-call cache of a per-environment safe action set (NavigateSeeker)
        self.fast = True                    # m = 1 boundary projection through the canonicalise-once path
        self._p1_prog = None
        self.m, self.n, self.N = env.action_dim, env.state_dim, env.num_envs
        # safeguard.py:54 -- linearised ONCE at construction; env 0's noise/action matrices are baked in (Q4)
        const, _, b_mat, e_mat = env.linear_dynamics()
        self.B0 = b_mat[0].detach().clone()
        self.G_w = (e_mat[0] @ torch.diag(env.noise_half)).detach().clone()
        self.interventions = 0
        self.safe_action = None
        self.infeasible = torch.zeros(self.N, dtype=torch.bool)
        self.projectable = torch.zeros(self.N, dtype=torch.bool)

    # -- parameters (safeguard.py:200-208) ------------------------------------------------------
    def _sets(self, i: int, const: Tensor, b_mat: Tensor) -> P.SafeSets:
        env = self.env
        s = P.SafeSets(m=self.m, n=self.n)
        if env.action_constrained:
            c_a, G_a = self._action_set if self._action_set is not None else env.safe_action_set()
            s.c_a, s.G_a = (c_a[i], G_a[i]) if G_a.dim() == 3 else (c_a, G_a)
        if env.state_constrained:
            s.c_s, s.G_s = env.safe_state_set()
            s.c_f, s.B, s.G_w = const[i], b_mat[i], self.G_w
        return s

    # -- Safeguard.actions (safeguard.py:61-80) -----------------------------------------------------
    def actions(self, action: Tensor, directions: Optional[Tensor] = None) -> Tensor:
        env = self.env
        center, gen = env.reachable_set()
        projectable = aab_intersects_zonotope(env.state_center, env.state_half, center, gen)
        self.projectable = projectable
        if self.passthrough:                 # ray_mask.py:21-30,48-50 (Q11): forward without graph, identity bwd
            with torch.no_grad():
                guarded = self.safeguard(action.detach(), directions)
                out = self._gate(projectable, action.detach(), guarded)
            safe_action = action + (out - action).detach()
        else:
            guarded = self.safeguard(action, directions)
            safe_action = self._gate(projectable, action, guarded)
        if self.strict and (safe_action.isnan().any() or safe_action.isinf().any()):
            raise ValueError("Safe action are NaN.")
        self.safe_action = safe_action
        differs = ~torch.isclose(safe_action, action)
        self.interventions += int((differs.count_nonzero(dim=1) == self.m).sum())
        return safe_action

    def _gate(self, projectable, action, guarded):
        if self.gate == "reference":         # Q1: raw action where projectable
            return torch.where(projectable.unsqueeze(1), action, guarded)
        if self.gate == "intended":
            return torch.where(projectable.unsqueeze(1), guarded, action)
        return guarded

    def step(self, action: Tensor, directions: Optional[Tensor] = None, noise: Optional[Tensor] = None):
        return self.env.step(self.actions(action, directions), noise)

    # -- the maps -------------------------------------------------------------------------------
    def safeguard(self, action: Tensor, directions: Optional[Tensor] = None) -> Tensor:
        self._action_set = self.env.safe_action_set() if self.env.env_name == "NavigateSeeker" else None
        if self.kind == "boundary_projection":
            return self.boundary_projection(action)
        return self.ray_mask(action, directions)

    def _linear(self):
        if self.env.state_constrained:
            const, _, b_mat, _ = self.env.linear_dynamics()      # second pass per step (Q14)
            return const, b_mat
        return None, None

    def boundary_projection(self, action: Tensor) -> Tensor:
        """boundary_projection.py:54-55: one P1 per environment."""
        const, b_mat = self._linear()
        if self.m == 1 and self.fast:
            # canonicalise once (like the CvxpyLayer built at the first call), refill (c_f, B) per environment
            if self._p1_prog is None:
                self._p1_prog = P.ParametricP1M1(self._sets(0, torch.zeros(self.N, self.n, dtype=F64),
                                                            torch.zeros(self.N, self.n, 1, dtype=F64))
                                                 if self.env.state_constrained else self._sets(0, None, None))
            theta = torch.cat([const, b_mat[:, :, 0]], dim=1) if self.env.state_constrained \
                else torch.zeros(self.N, 0, dtype=F64)
            a_s, infeasible = P.p1_projection_batch_m1(action, self._p1_prog, theta)
            self.infeasible = infeasible
            return a_s
        out = []
        self.infeasible = torch.zeros(self.N, dtype=torch.bool)
        for i in range(self.N):
            try:
                a_s = P.p1_projection(action[i], self._sets(i, const, b_mat))[0]
            except P.Infeasible:
                self.infeasible[i] = True
                a_s = torch.full((self.m,), float("nan"), dtype=F64)
            out.append(a_s)
        return torch.stack(out)

    # ray_mask.py:63-89
    def ray_mask(self, action: Tensor, directions: Optional[Tensor]) -> Tensor:
        env = self.env
        if env.state_constrained:
            if self.zonotopic:
                center, safe_dist, feas_dist = self.zonotopic_approximation(action, directions)
            else:
                center, safe_dist, feas_dist = self.orthogonal_approximation(action)
        else:
            c_a, g_a = self._action_set if self._action_set is not None else env.safe_action_set()
            center = c_a if c_a.dim() == 2 else c_a.expand(self.N, -1)
            safe_dist, feas_dist = self.compute_distances(action, center, g_a if g_a.dim() == 3 else g_a.expand(self.N, -1, -1))
        action_dist = torch.linalg.vector_norm(action - center, dim=1, ord=2, keepdim=True)
        dirs = (action - center) / (action_dist + 1e-8)
        central = action_dist < 1e-8
        return torch.where(central, center, center + dirs * self.radial_mapping(action_dist, safe_dist, feas_dist))

    def radial_mapping(self, action_dist, safe_dist, feas_dist):
        """ray_mask.py:108-113, as shipped (Q2)."""
        if self.linear_projection:
            return action_dist / safe_dist
        return torch.tanh(action_dist / safe_dist) / torch.tanh(feas_dist / safe_dist)

    def draw_directions(self) -> Tensor:
        """ray_mask.py:184-185 (Q9: consumes the global RNG inside the safeguard)."""
        d = self.env.rng("rand", (self.N, self.m, 2 * self.m)) * 2 - 1
        return d / torch.linalg.vector_norm(d, dim=1, keepdim=True)

    def zonotopic_approximation(self, action, directions):
        """ray_mask.py:130-132 -> zonotope_expansion (P2) + compute_distances (P3)."""
        if directions is None:
            directions = self.draw_directions()
        const, b_mat = self._linear()
        centers, gens = [], []
        self.infeasible = torch.zeros(self.N, dtype=torch.bool)
        for i in range(self.N):
            try:
                c, length = P.p2_expansion(directions[i], self._sets(i, const, b_mat), self.B0)
            except P.Infeasible:
                self.infeasible[i] = True
                c = torch.full((self.m,), float("nan"), dtype=F64)
                length = torch.full((2 * self.m,), float("nan"), dtype=F64)
            centers.append(c)
            gens.append(directions[i] * length.unsqueeze(0))       # ray_mask.py:189
        center, gen = torch.stack(centers), torch.stack(gens)
        safe_dist, feas_dist = self.compute_distances(action, center, gen)
        return center, safe_dist, feas_dist

    def compute_distances(self, action, center, gen):
        """ray_mask.py:236-240."""
        diff = action - center
        dirs = diff / (torch.linalg.vector_norm(diff, dim=1, keepdim=True) + 1e-8)
        dist = []
        for i in range(self.N):
            if self.infeasible[i]:
                dist.append(torch.full((), float("nan"), dtype=F64))
                continue
            dist.append(P.p3_ray_distance(dirs[i], center[i], gen[i]).reshape(()))
        safe_dist = torch.stack(dist).unsqueeze(1)
        return safe_dist, self.unit_box_dist(center, dirs)

    @staticmethod
    def unit_box_dist(center: Tensor, direction: Tensor) -> Tensor:
        """ray_mask.py:243-266: distance from ``center`` to the boundary of [-1,1]^m along ``direction``."""
        inf = torch.full_like(center, float("inf"))
        nz = direction != 0
        safe_dir = torch.where(nz, direction, torch.ones_like(direction))
        hi = torch.where(nz, (1 - center) / safe_dir, inf)
        lo = torch.where(nz, (-1 - center) / safe_dir, inf)
        both = torch.cat([hi, lo], dim=-1)
        return torch.min(torch.where(both < 0, torch.full_like(both, float("inf")), both), dim=-1, keepdim=True).values

    def orthogonal_approximation(self, action):
        """ray_mask.py:268-341: P1 start point, P4 chord, midpoint centre."""
        const, b_mat = self._linear()
        start = self.boundary_projection(action)
        action_dist = torch.linalg.vector_norm(start - action, dim=1, ord=2, keepdim=True)
        safe = action_dist < 1e-8
        # the reference divides by action_dist itself (ray_mask.py:329); with an exact P1 the distance of an
        # already-safe action is exactly 0 and 0/0 would poison the gradient, so the masked rows divide by 1
        dirs = (start - action) / torch.where(safe, torch.ones_like(action_dist), action_dist)
        dist = []
        for i in range(self.N):
            if bool(safe[i]) or self.infeasible[i]:
                # the reference still solves P4 with a NaN direction here and masks the result; value unused
                dist.append(torch.ones((), dtype=F64))
                continue
            dist.append(P.p4_implicit_distance(start[i], dirs[i], self._sets(i, const, b_mat)).reshape(()))
        dist = torch.stack(dist).unsqueeze(1)
        safe_dist = torch.where(safe, torch.ones_like(dist), dist / 2)
        dirs_clean = torch.where(safe, torch.zeros_like(dirs), dirs)
        center = torch.where(safe, start, start + dirs_clean * safe_dist)
        feas = torch.where(safe, torch.ones_like(safe_dist), self.unit_box_dist(center, -dirs_clean))
        return center, safe_dist, feas
