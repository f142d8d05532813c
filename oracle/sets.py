"""Oracle set algebra (float64, CPU).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates the pieces of src/sets the hot path touches: the separating-axis gate
``Box.intersects(Zonotope)`` (box.py:139-154), ``Box.contains`` (box.py:77-108),
``Zonotope.box`` (zonotope.py:267-276) and the verdict programs P9
``Zonotope.contains`` (zonotope.py:143-239), the latter solved with scipy-HiGHS.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.optimize import linprog
from torch import Tensor


def box_edges(generator: Tensor):
    """Box.__init__ (box.py:33-43): edge lengths/directions of an axis-aligned (diagonal) box."""
    edge_len = torch.linalg.vector_norm(generator, dim=1)          # column norms
    edge_dir = generator / edge_len.unsqueeze(1)
    return edge_len, edge_dir


def aab_intersects_zonotope(box_center: Tensor, box_half: Tensor, z_center: Tensor, z_gen: Tensor) -> Tensor:
    """``state_set.intersects(reachable_set)`` -- box.py:139-154, the over-approximating gate (Q1).

    box_center/box_half: (n,) ; z_center: (N,n) ; z_gen: (N,n,p).  Returns bool (N,).
    Two families of axes: the box edges, and the *normalised zonotope generators* themselves.
    """
    n = box_center.shape[0]
    N = z_center.shape[0]
    box_gen = torch.diag(box_half).expand(N, n, n)
    edge_len = box_half.expand(N, n)
    edge_dir = torch.eye(n, dtype=z_center.dtype).expand(N, n, n)
    center = z_center - box_center
    proj = (center.unsqueeze(2) * edge_dir).sum(dim=1).abs()
    supports = (edge_dir.unsqueeze(2) * z_gen.unsqueeze(3)).sum(dim=1).abs().sum(dim=1)
    o_len = torch.linalg.vector_norm(z_gen, dim=1)
    o_dir = z_gen / o_len.unsqueeze(1)
    o_proj = (center.unsqueeze(2) * o_dir).sum(dim=1).abs()
    o_supports = (o_dir.unsqueeze(2) * box_gen.unsqueeze(3)).sum(dim=1).abs().sum(dim=1)
    return (proj <= edge_len + supports).all(dim=1) & (o_proj <= o_len + o_supports).all(dim=1)


def aab_contains_point(box_center: Tensor, box_half: Tensor, p: Tensor) -> Tensor:
    """Box.contains(Tensor) for an axis-aligned box (box.py:90-93)."""
    return ((p - box_center).abs() <= box_half).all(dim=1)


def aab_contains_zonotope(box_center: Tensor, box_half: Tensor, z_center: Tensor, z_gen: Tensor) -> Tensor:
    """Box.contains(Zonotope) (box.py:98-102)."""
    return ((z_center - box_center).abs() + z_gen.abs().sum(dim=2) <= box_half).all(dim=1)


def zonotope_point_norm(center: np.ndarray, gen: np.ndarray, point: np.ndarray) -> float:
    """P9a, zonotope.py:157-172:  min ||gamma||_inf  s.t.  G gamma = p - c.  Verdict = value <= 1.

    Returns +inf when the point is outside the generators' span.
    """
    d, p = gen.shape
    # variables [gamma (p), t]; minimise t; -t <= gamma_j <= t
    cost = np.zeros(p + 1)
    cost[-1] = 1.0
    a_ub = np.zeros((2 * p, p + 1))
    for j in range(p):
        a_ub[2 * j, j], a_ub[2 * j, -1] = 1.0, -1.0
        a_ub[2 * j + 1, j], a_ub[2 * j + 1, -1] = -1.0, -1.0
    a_eq = np.hstack([gen, np.zeros((d, 1))])
    res = linprog(cost, A_ub=a_ub, b_ub=np.zeros(2 * p), A_eq=a_eq, b_eq=point - center,
                  bounds=[(None, None)] * p + [(0, None)], method="highs")
    return float(res.fun) if res.status == 0 else float("inf")


def zonotope_contains_point(center, gen, point) -> bool:
    return zonotope_point_norm(np.asarray(center, float), np.asarray(gen, float), np.asarray(point, float)) <= 1.0


def zonotope_in_zonotope_norm(c1, g1, c2, g2) -> float:
    """P9b, zonotope.py:213-236 (Sadraddini-Tedrake sufficient encoding):
    min || [Gamma, beta] ||_inf(row-sum)  s.t.  G1 = G2 Gamma,  c2 - c1 = G2 beta.   Verdict = value <= 1.
    """
    c1, g1, c2, g2 = (np.asarray(x, float) for x in (c1, g1, c2, g2))
    d, q1 = g1.shape
    p2 = g2.shape[1]
    nv = p2 * (q1 + 1)              # Gamma (p2 x q1) row-major, then beta (p2)
    # variables: x (nv), abs-bounds u (nv), t
    tot = 2 * nv + 1
    cost = np.zeros(tot)
    cost[-1] = 1.0
    rows, rhs = [], []
    for k in range(nv):             # |x_k| <= u_k
        r = np.zeros(tot); r[k], r[nv + k] = 1.0, -1.0; rows.append(r); rhs.append(0.0)
        r = np.zeros(tot); r[k], r[nv + k] = -1.0, -1.0; rows.append(r); rhs.append(0.0)
    for i in range(p2):             # row sums <= t
        r = np.zeros(tot)
        for j in range(q1):
            r[nv + i * q1 + j] = 1.0
        r[nv + p2 * q1 + i] = 1.0
        r[-1] = -1.0
        rows.append(r); rhs.append(0.0)
    eq_rows, eq_rhs = [], []
    for a in range(d):
        for j in range(q1):         # (G2 Gamma)[a, j] = G1[a, j]
            r = np.zeros(tot)
            for i in range(p2):
                r[i * q1 + j] = g2[a, i]
            eq_rows.append(r); eq_rhs.append(g1[a, j])
        r = np.zeros(tot)           # (G2 beta)[a] = (c2 - c1)[a]
        for i in range(p2):
            r[p2 * q1 + i] = g2[a, i]
        eq_rows.append(r); eq_rhs.append(c2[a] - c1[a])
    res = linprog(cost, A_ub=np.array(rows), b_ub=np.array(rhs), A_eq=np.array(eq_rows), b_eq=np.array(eq_rhs),
                  bounds=[(None, None)] * nv + [(0, None)] * (nv + 1), method="highs")
    return float(res.fun) if res.status == 0 else float("inf")


def zonotope_contains_zonotope(c1, g1, c2, g2) -> bool:
    """Is Z1=<c1,g1> inside Z2=<c2,g2> according to the reference's encoding."""
    return zonotope_in_zonotope_norm(c1, g1, c2, g2) <= 1.0
