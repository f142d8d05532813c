"""Zonotope container (reference: src/sets/zonotope.py:15-80, 143-276).

``contains`` returns the reference's verdicts: the point / zonotope-in-zonotope LPs of
zonotope.py:157-172, 213-236 (value <= 1).  They are evaluated per batch element on the host with
scipy-HiGHS exactly like the reference evaluates them on the host with Clarabel (Q8: tests and
diagnostics only -- the per-step verdict bits of the rollout come from the CUDA kernels)."""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from .. import rng
from .interface import ConvexSet


class Zonotope(ConvexSet):
    def __init__(self, center: Tensor, generator: Tensor):
        if center.dim() != 2 or generator.dim() != 3 or generator.shape[:2] != center.shape:
            raise TypeError(f"Zonotope: center {tuple(center.shape)} / generator {tuple(generator.shape)}")
        super().__init__(*center.shape)
        self.center, self.generator = center, generator
        self.num_gens = generator.shape[2]

    def __iter__(self):
        return iter((self.center, self.generator))

    def __getitem__(self, item):
        if isinstance(item, int):
            return Zonotope(self.center[item].unsqueeze(0), self.generator[item].unsqueeze(0))
        if isinstance(item, Tensor):
            return Zonotope(self.center[item], self.generator[item])
        raise TypeError(f"Invalid argument type {type(item)}")

    def sample(self) -> Tensor:
        return self.center + torch.sum(self.generator * (2 * rng.rand(self.batch_dim, 1, self.num_gens) - 1), dim=2)

    def box(self):
        from .box import Box
        return Box(self.center, self.generator.abs().sum(dim=2).diag_embed())

    def contains(self, other) -> Tensor:
        from scipy.optimize import linprog
        out = []
        for b in range(self.batch_dim):
            c = self.center[b].detach().cpu().numpy().astype(float)
            G = self.generator[b].detach().cpu().numpy().astype(float)
            if isinstance(other, Tensor):
                out.append(_point_norm(c, G, other[b].detach().cpu().numpy().astype(float), linprog) <= 1.0)
            elif isinstance(other, Zonotope):
                out.append(_zono_norm(other.center[b].detach().cpu().numpy().astype(float),
                                      other.generator[b].detach().cpu().numpy().astype(float), c, G, linprog) <= 1.0)
            else:
                raise NotImplementedError(f"Containment check not implemented for {type(other)}")
        return torch.tensor(out, dtype=torch.bool, device=self.center.device)

    def intersects(self, other) -> Tensor:
        from .box import Box
        if isinstance(other, Box):
            return other.intersects(self)
        if isinstance(other, Zonotope):
            return self.box().intersects(other.box())       # over-approximation, zonotope.py:260-262
        return other.intersects(self)


def _point_norm(c, G, p, linprog):
    d, k = G.shape
    cost = np.zeros(k + 1); cost[-1] = 1.0
    A = np.zeros((2 * k, k + 1))
    for j in range(k):
        A[2 * j, j], A[2 * j, -1] = 1.0, -1.0
        A[2 * j + 1, j], A[2 * j + 1, -1] = -1.0, -1.0
    res = linprog(cost, A_ub=A, b_ub=np.zeros(2 * k), A_eq=np.hstack([G, np.zeros((d, 1))]), b_eq=p - c,
                  bounds=[(None, None)] * k + [(0, None)], method="highs")
    return res.fun if res.status == 0 else np.inf


def _zono_norm(c1, G1, c2, G2, linprog):
    d, q1 = G1.shape
    p2 = G2.shape[1]
    nx = p2 * (q1 + 1)
    tot = 2 * nx + 1
    cost = np.zeros(tot); cost[-1] = 1.0
    ub = []
    for i in range(nx):
        r = np.zeros(tot); r[i], r[nx + i] = 1.0, -1.0; ub.append(r)
        r = np.zeros(tot); r[i], r[nx + i] = -1.0, -1.0; ub.append(r)
    for i in range(p2):
        r = np.zeros(tot); r[nx + i * (q1 + 1): nx + (i + 1) * (q1 + 1)] = 1.0; r[-1] = -1.0; ub.append(r)
    eq, beq = [], []
    for a in range(d):
        for j in range(q1 + 1):
            r = np.zeros(tot)
            for i in range(p2):
                r[i * (q1 + 1) + j] = G2[a, i]
            eq.append(r)
            beq.append(G1[a, j] if j < q1 else c2[a] - c1[a])
    res = linprog(cost, A_ub=np.array(ub), b_ub=np.zeros(len(ub)), A_eq=np.array(eq), b_eq=np.array(beq),
                  bounds=[(None, None)] * nx + [(0, None)] * (nx + 1), method="highs")
    return res.fun if res.status == 0 else np.inf
