"""Constant-registration time reduction of the safe sets (host side, float64, runs once per wrapper).

The reference canonicalises its cvxpy problems once, at the first safeguard call, baking the constant
matrices in as numpy arrays (safeguard.py:111-122; ray_mask.py:157-171; Q4).  The B200 build does the
analogous one-off work here: it eliminates the lifted variables (gamma, Gamma, beta) of the containment
encodings (sets/zonotope.py:82-141) so that the kernels only see rows in ACTION space:

* safe action zonotope  ->  interval (m = 1) or polygon half-planes (m = 2);
* safe state zonotope with invertible G_s  ->  G_s^-1, row budgets 1 - sum_j |(G_s^-1 G_w)_ij|;
* safe state zonotope with n = 2 and p > 2 generators (pendulum RCI set)  ->  the exact H-representation
  of the feasible-offset polygon K and of the polytope Q of P2, found with a support-function
  refinement whose LPs run on scipy-HiGHS (init-time only; no solver on the per-step path).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from ._lib import MAXHP, MAXM, MAXN, Consts

BOXES = {   # feasible state box (centre, half) and noise centre per simulator; pendulum.py:65-72 etc.
    "pendulum": ([0, 0], [np.pi, 8.0], [0, 0], [0.0, 0.1]),
    "cartpole": ([0] * 4, [10.0, np.pi, 10.0, 10.0], [0] * 4, [0, 0, 0, 0.1]),
    "quadrotor": ([0] * 6, [8.0, 8.0, np.pi / 12, 0.8, 1.0, np.pi / 2], [0] * 6, [0, 0, 0, 0.1, 0.1, 0]),
    "seeker": ([0, 0], [8.0, 8.0], [0, 0], [0.05, 0.05]),
    "household": ([5.0, 21.0, 55.0], [5.0, 3.0, 45.0], [0.0, 15.0, 0.0], [0.0, 25.0, 0.0]),
}
ENV_SIM = {"BalancePendulum": "pendulum", "BalanceCartPole": "cartpole", "SwingUpCartPole": "cartpole",
           "BalanceQuadrotor": "quadrotor", "ManageHousehold": "household", "NavigateSeeker": "seeker",
           "NavigateQuadrotor": "quadrotor"}


def zonotope_polygon_rows(center: np.ndarray, gen: np.ndarray) -> np.ndarray:
    """H-representation of a 2-D zonotope: one pair of rows per generator direction.  (rows, 3): n0 n1 h."""
    rows = []
    seen = []
    for j in range(gen.shape[1]):
        g = gen[:, j]
        nrm = np.linalg.norm(g)
        if nrm == 0.0:
            continue
        n = np.array([-g[1], g[0]]) / nrm
        if any(abs(abs(n @ s) - 1.0) < 1e-13 for s in seen):
            continue
        seen.append(n)
        h = np.abs(n @ gen).sum()
        rows.append([n[0], n[1], h + n @ center])
        rows.append([-n[0], -n[1], h - n @ center])
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------
# polygon model: exact facets of K (and Q) from a support-function oracle
# ------------------------------------------------------------------------------------------------
def _containment_lp(G_s: np.ndarray, fixed_cols: np.ndarray, free_col: Optional[np.ndarray]):
    """Variables [Gamma (p x q), beta (p), (L), U (p x (q+1))] of  enc(Z(., [free_col*L, fixed_cols]) in Z(., G_s)).

    Returns (A_ub, b_ub, A_eq, b_eq, idx_beta, idx_L) with d = G_s beta the free offset.
    """
    n, p = G_s.shape
    cols = fixed_cols.shape[1] + (1 if free_col is not None else 0)
    nG, nb = p * cols, p
    iL = nG + nb if free_col is not None else None
    nU0 = nG + nb + (1 if free_col is not None else 0)
    nv = nU0 + p * (cols + 1)
    eq, beq, ub, bub = [], [], [], []
    for j in range(cols):
        for k in range(n):
            r = np.zeros(nv)
            for i in range(p):
                r[i * cols + j] = G_s[k, i]
            if free_col is not None and j == 0:
                r[iL] = -free_col[k]
                beq.append(0.0)
            else:
                beq.append(fixed_cols[k, j - (1 if free_col is not None else 0)])
            eq.append(r)
    for i in range(p):
        for j in range(cols + 1):
            x = i * cols + j if j < cols else nG + i
            u = nU0 + i * (cols + 1) + j
            r = np.zeros(nv); r[x], r[u] = 1.0, -1.0; ub.append(r); bub.append(0.0)
            r = np.zeros(nv); r[x], r[u] = -1.0, -1.0; ub.append(r); bub.append(0.0)
        r = np.zeros(nv)
        r[nU0 + i * (cols + 1): nU0 + (i + 1) * (cols + 1)] = 1.0
        ub.append(r); bub.append(1.0)
    return np.array(ub), np.array(bub), np.array(eq), np.array(beq), slice(nG, nG + nb), iL


def _support_points(G_s, fixed_cols, free_col, directions):
    from scipy.optimize import linprog   # init-time only
    A_ub, b_ub, A_eq, b_eq, ib, iL = _containment_lp(G_s, fixed_cols, free_col)
    nv = A_ub.shape[1]
    pts = []
    for u in directions:
        cost = np.zeros(nv)
        cost[ib] = -(u[:G_s.shape[0]] @ G_s)
        if iL is not None:
            cost[iL] = -u[-1]
        res = linprog(cost, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq if len(A_eq) else None, b_eq=b_eq if len(b_eq) else None,
                      bounds=[(None, None)] * nv, method="highs")
        if res.status != 0:
            raise ValueError(f"safe-state reduction LP failed: {res.message}")
        d = G_s @ res.x[ib]
        pts.append(np.concatenate([d, [res.x[iL]]]) if iL is not None else d)
    return np.array(pts)


def merge_parallel_generators(G: np.ndarray, tol: float = 1e-12) -> np.ndarray:
    """Drop zero generators and merge those of equal direction (their lengths add): same zonotope, and the same
    containment encoding (rows with equal generators can share one budget by convexity)."""
    cols = []
    for j in range(G.shape[1]):
        g = G[:, j]
        nrm = np.linalg.norm(g)
        if nrm < tol:
            continue
        for c in cols:
            cn = np.linalg.norm(c)
            cosang = (c @ g) / (cn * nrm)
            if abs(abs(cosang) - 1.0) < 1e-12:
                c += np.sign(cosang) * g
                break
        else:
            cols.append(g.astype(np.float64).copy())
    return np.stack(cols, axis=1)


def polytope_facets(G_s: np.ndarray, fixed_cols: np.ndarray, free_col: Optional[np.ndarray] = None,
                    tol: float = 1e-10) -> np.ndarray:
    """Exact facets  h . x <= 1  of  {x = (d[, L]) : the containment encoding is feasible}.

    Beneath-beyond with a support oracle: grow the vertex set until every hull facet is supported.
    """
    from scipy.spatial import ConvexHull   # init-time only
    dim = G_s.shape[0] + (1 if free_col is not None else 0)
    dirs = [s * e for e in np.eye(dim) for s in (1.0, -1.0)]
    rng = np.random.default_rng(0)
    dirs += list(rng.normal(size=(4 * dim, dim)))
    pts = _support_points(G_s, fixed_cols, free_col, dirs)
    for _ in range(200):
        pts = np.unique(np.round(pts, 12), axis=0)
        hull = ConvexHull(pts)
        eqs = hull.equations                      # n . x + off <= 0
        normals, offs = eqs[:, :-1], -eqs[:, -1]
        uniq = {}
        for nrm, off in zip(normals, offs):
            uniq[tuple(np.round(np.append(nrm, off), 9))] = (nrm, off)
        cand = _support_points(G_s, fixed_cols, free_col, [v[0] for v in uniq.values()])
        new = [c for c, (nrm, off) in zip(cand, uniq.values()) if nrm @ c > off + tol * (1 + abs(off))]
        if not new:
            rows = np.array([nrm / off for nrm, off in uniq.values()])
            return rows
        pts = np.vstack([pts, np.array(new)])
    raise ValueError("polytope facet enumeration did not converge")


# ------------------------------------------------------------------------------------------------
def build_consts(env_name: str, sg_kind: int, *, gate_mode: int = 0, rm_linear: bool = True,
                 rm_zonotopic: bool = True, rm_passthrough: bool = False,
                 safe_action: Optional[tuple] = None, safe_state: Optional[tuple] = None,
                 G_w: Optional[np.ndarray] = None, B0: Optional[np.ndarray] = None,
                 dynamic_action_set: bool = False, n_obstacles: int = 0, dynamic_state_set: bool = False) -> Consts:
    """Fill ``struct sgbpo_consts``.  safe_action / safe_state: (center, generator) numpy float64 or None."""
    sim = ENV_SIM[env_name]
    bc, bh, nc, _ = (np.asarray(x, dtype=np.float64) for x in BOXES[sim])
    n = bc.shape[0]
    k = Consts()
    k.sg_kind, k.gate_mode = int(sg_kind), int(gate_mode)
    k.rm_linear, k.rm_zonotopic, k.rm_passthrough = int(rm_linear), int(rm_zonotopic), int(rm_passthrough)
    for i in range(n):
        k.box_min[i], k.box_max[i], k.noise_center[i] = bc[i] - bh[i], bc[i] + bh[i], nc[i]
    k.action_constrained = int(safe_action is not None or dynamic_action_set)
    k.n_obstacles = int(n_obstacles)
    k.act_set_mode = int(dynamic_action_set)
    k.state_constrained = int(safe_state is not None or dynamic_state_set)
    if dynamic_state_set:
        # NavigateQuadrotor: the safe state set is rebuilt per environment every step (csrc/sgbpo_navquad.cuh, P6-P8); the only
        # registered constant is the noise generator G_w = E_0 G_noise of the init-time linearisation (Q4), stored in Gam_p
        k.state_model = 4
        Gw = np.zeros((n, n)) if G_w is None else np.asarray(G_w, dtype=np.float64)
        Wn = Gw[:, np.abs(Gw).sum(axis=0) > 0]
        if Wn.shape[1] > 2:
            raise ValueError("NavigateQuadrotor: at most two noise generators")
        k.n_wcol = Wn.shape[1]
        for i in range(n):
            for j in range(Wn.shape[1]):
                k.Gam_p[i][j] = Wn[i, j]
    if safe_action is not None:
        c_a, G_a = (np.asarray(x, dtype=np.float64) for x in safe_action)
        m = c_a.shape[0]
        for q in range(m):
            k.c_a[q] = c_a[q]
        if m == 1:
            r = np.abs(G_a).sum()
            k.a_lo[0], k.a_hi[0] = c_a[0] - r, c_a[0] + r
        else:
            rows = zonotope_polygon_rows(c_a, G_a)
            if len(rows) > MAXHP:
                raise ValueError(f"safe action polygon has {len(rows)} rows > {MAXHP}")
            k.n_act_hp = len(rows)
            for r, row in enumerate(rows):
                for j in range(3):
                    k.act_hp[r][j] = row[j]
    if safe_state is not None:
        c_s, G_s = (np.asarray(x, dtype=np.float64) for x in safe_state)
        G_w = np.zeros((n, n)) if G_w is None else np.asarray(G_w, dtype=np.float64)
        m = 1 if B0 is None else np.asarray(B0).shape[1]
        B0 = np.zeros((n, m)) if B0 is None else np.asarray(B0, dtype=np.float64)
        for i in range(n):
            k.c_s[i] = c_s[i]
        if G_s.shape[0] == G_s.shape[1] and abs(np.linalg.det(G_s)) > 1e-12:
            k.state_model = 0
            Ginv = np.linalg.inv(G_s)
            cs_hat, rb, B0h = Ginv @ c_s, 1.0 - np.abs(Ginv @ G_w).sum(axis=1), np.abs(Ginv @ B0)
            for i in range(n):
                k.cs_hat[i], k.rb[i] = cs_hat[i], rb[i]
                for j in range(n):
                    k.Ginv[i * n + j] = Ginv[i, j]
                for q in range(m):
                    k.B0hat_abs[i * MAXM + q] = B0h[i, q]
                    k.B0hat[i * MAXM + q] = (Ginv @ B0)[i, q]
        elif n == 2 and m == 1:
            k.state_model = 1
            nz = G_w[:, np.abs(G_w).sum(axis=0) > 0]
            K_rows = polytope_facets(G_s, nz)
            if len(K_rows) > MAXHP:
                raise ValueError(f"state polygon K has {len(K_rows)} rows > {MAXHP}")
            k.n_k_hp = len(K_rows)
            for r, row in enumerate(K_rows):
                k.k_hp[r][0], k.k_hp[r][1] = row[0], row[1]
            if sg_kind == _lib.SG_RAY_MASK and rm_zonotopic:
                Q_rows = polytope_facets(G_s, nz, free_col=B0[:, 0])
                Q_rows = Q_rows[Q_rows[:, 2] > -1e-12]         # Q is symmetric in L: for L >= 0 the rows with a
                                                               # negative L coefficient are implied by their mirror
                if len(Q_rows) > MAXHP:
                    raise ValueError(f"state polytope Q has {len(Q_rows)} rows > {MAXHP}")
                k.n_q_hp = len(Q_rows)
                for r, row in enumerate(Q_rows):
                    for j in range(3):
                        k.q_hp[r][j] = row[j]
        else:
            # G_s wide (BalanceQuadrotor: 6 x 24): the feasible offsets form a 6-D polytope whose facets cannot be
            # tabulated (beneath-beyond on it does not terminate in reasonable time), so the containment encoding
            # stays in its lifted form and every environment solves it (interior point, csrc/sgbpo_lifted.cuh).
            # Generators of equal direction are merged and zero ones dropped (equivalent encoding: the merged row
            # budget is the mean of the merged rows' budgets, convexity), which leaves 14 of the 24 columns.
            G = merge_parallel_generators(G_s)
            Wn = G_w[:, np.abs(G_w).sum(axis=0) > 0]
            p, nn = G.shape[1], G.shape[1] - np.linalg.matrix_rank(G)
            built = sg_kind == _lib.SG_BOUNDARY_PROJECTION or (sg_kind == _lib.SG_RAY_MASK and not rm_zonotopic)
            if not built or p > _lib.MAXGEN or nn > _lib.MAXNULL or Wn.shape[1] > 2 or m != 2 \
                    or np.linalg.matrix_rank(G) < n:
                k.state_model = 2          # unreduced: gate + raw action only (zonotopic ray mask P2 of this set: not built)
            else:
                k.state_model = 3
                Gp = np.linalg.pinv(G)
                from scipy.linalg import null_space
                Nm = null_space(G)
                k.n_gen, k.n_null, k.n_wcol = p, Nm.shape[1], Wn.shape[1]
                Gam = Gp @ Wn
                for i in range(p):
                    for j in range(n):
                        k.Gpinv[i][j] = Gp[i, j]
                    for j in range(Nm.shape[1]):
                        k.Nmat[i][j] = Nm[i, j]
                    for j in range(Wn.shape[1]):
                        k.Gam_p[i][j] = Gam[i, j]
    return k
