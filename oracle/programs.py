"""Oracle convex programs P1-P4 in their generic lifted form.  TEST INFRASTRUCTURE (oracle/__init__.py).

Third-party arithmetic this replaces (absent from /root/reference): cvxpy>=1.7.1 / cvxpylayers>=0.1.9
-> diffcp -> Clarabel (pyproject.toml:8-9; safeguard.py:27).  Published algorithm restated: the
reference hands cvxpy a *disciplined parametrised program*; the layer returns the optimal primal
variables and, in backward, the derivative of the optimal solution w.r.t. every parameter
(implicit function theorem on the KKT / homogeneous self-dual embedding).  Away from degenerate
points that derivative equals the derivative of the solution of the KKT system restricted to the
optimal active set, which is what this file computes.

Formulation (all variables explicit, no per-environment shortcut):
  pt(p in Z(c,G))         : p - c = G gamma, |gamma_j| <= 1                       zonotope.py:102-106
  enc(Z(c1,G1) in Z(c2,G2)): G1 = G2 Gamma, c2 - c1 = G2 beta,
                            max_i( sum_j |Gamma_ij| + |beta_i| ) <= 1             zonotope.py:131-141
  linear step              : c+(a) = c_f + B (a - a_lin),  G+ = E0 G_noise        safeguard.py:98-122
  feasibility              : -1 <= a <= 1                                        safeguard.py:125-139

Solvers: scipy-HiGHS for every LP; gradients by the Lagrangian (envelope) identity for optimal
values and by a torch pseudo-inverse solve of the active KKT system for optimal points, so that
``torch.autograd`` carries them exactly like the reference's CvxpyLayer would.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from scipy.optimize import linprog, minimize
from torch import Tensor

F64 = torch.float64
ACTIVE_TOL = 1e-9


class Infeasible(Exception):
    """Raised where the reference's CvxpyLayer would raise a solver error (Q13)."""


# ------------------------------------------------------------------------------------------------
# a tiny differentiable LP container
# ------------------------------------------------------------------------------------------------
class Model:
    """min c^T x  s.t.  A_eq x = b_eq,  A_ub x <= b_ub.  Rows are torch tensors (carry autograd)."""

    def __init__(self):
        self.nvar = 0
        self.blocks = {}
        self.eq, self.ub = [], []       # lists of (row tensor (nvar,), rhs tensor ())

    def var(self, name: str, size: int) -> slice:
        sl = slice(self.nvar, self.nvar + size)
        self.blocks[name] = sl
        self.nvar += size
        return sl

    def add_eq(self, terms, rhs):
        self.eq.append((terms, rhs))

    def add_ub(self, terms, rhs):
        self.ub.append((terms, rhs))

    def finalize(self):
        """Assemble dense A, b with one index_put per block (coefficients may carry autograd)."""
        def stack(rows):
            if not rows:
                return torch.zeros(0, self.nvar, dtype=F64), torch.zeros(0, dtype=F64)
            ri, ci, vals = [], [], []
            for r, (terms, _) in enumerate(rows):
                for idx, coef in terms:
                    ri.append(r)
                    ci.append(idx)
                    vals.append(coef if isinstance(coef, Tensor) else torch.tensor(float(coef), dtype=F64))
            a = torch.zeros(len(rows), self.nvar, dtype=F64)
            if vals:
                a = a.index_put((torch.tensor(ri), torch.tensor(ci)),
                                torch.stack([v.reshape(()).to(F64) for v in vals]), accumulate=True)
            b = torch.stack([(r if isinstance(r, Tensor) else torch.tensor(float(r), dtype=F64)).reshape(()).to(F64)
                             for _, r in rows])
            return a, b
        self.A_eq, self.b_eq = stack(self.eq)
        self.A_ub, self.b_ub = stack(self.ub)
        return self


def _abs_rows(model: Model, x_idx: int, u_idx: int):
    """|x| <= u  as two rows."""
    model.add_ub([(x_idx, 1.0), (u_idx, -1.0)], 0.0)
    model.add_ub([(x_idx, -1.0), (u_idx, -1.0)], 0.0)


def add_point_containment(model: Model, tag: str, point_terms, point_const, center: Tensor, gen: Tensor):
    """pt(p in Z(center, gen)) where p_k = point_const[k] + sum(point_terms[k]) is affine in the variables.

    zonotope.py:102-106: ``point - center == generator @ weights``, ``norm(weights, inf) <= 1``.
    """
    d, p = gen.shape
    w = model.var(f"{tag}_gamma", p)
    for k in range(d):
        terms = list(point_terms[k]) + [(w.start + j, -gen[k, j]) for j in range(p)]
        model.add_eq(terms, center[k] - point_const[k])
    for j in range(p):
        model.add_ub([(w.start + j, 1.0)], 1.0)
        model.add_ub([(w.start + j, -1.0)], 1.0)
    return w


def add_zonotope_containment(model: Model, tag: str, c1_terms, c1_const, g1_cols, c2: Tensor, g2: Tensor):
    """enc(Z(c1,G1) in Z(c2,G2)), zonotope.py:131-141.

    c1_k = c1_const[k] + sum(c1_terms[k]) (affine in variables);
    g1_cols: list over columns j of (terms_per_row, const_per_row): G1[k,j] affine in variables.
    """
    d, p2 = g2.shape
    q1 = len(g1_cols)
    gam = model.var(f"{tag}_Gamma", p2 * q1)          # row-major (i, j)
    beta = model.var(f"{tag}_beta", p2)
    u = model.var(f"{tag}_U", p2 * (q1 + 1))
    for j, (col_terms, col_const) in enumerate(g1_cols):     # shape: G1 = G2 Gamma
        for k in range(d):
            terms = [(gam.start + i * q1 + j, g2[k, i]) for i in range(p2)]
            terms += [(idx, -coef) for idx, coef in col_terms[k]]
            model.add_eq(terms, col_const[k])
    for k in range(d):                                       # position: c2 - c1 = G2 beta
        terms = [(beta.start + i, g2[k, i]) for i in range(p2)]
        terms += list(c1_terms[k])
        model.add_eq(terms, c2[k] - c1_const[k])
    for i in range(p2):                                      # size: row sums of |[Gamma beta]| <= 1
        for j in range(q1):
            _abs_rows(model, gam.start + i * q1 + j, u.start + i * (q1 + 1) + j)
        _abs_rows(model, beta.start + i, u.start + i * (q1 + 1) + q1)
        model.add_ub([(u.start + i * (q1 + 1) + j, 1.0) for j in range(q1 + 1)], 1.0)
    return gam, beta


# ------------------------------------------------------------------------------------------------
# solving + differentiable values
# ------------------------------------------------------------------------------------------------
def solve_lp(model: Model, cost: Tensor):
    """HiGHS solve.  Returns (x (numpy), value_t (torch, differentiable via the Lagrangian identity))."""
    res = linprog(cost.detach().numpy(),
                  A_ub=model.A_ub.detach().numpy() if len(model.ub) else None,
                  b_ub=model.b_ub.detach().numpy() if len(model.ub) else None,
                  A_eq=model.A_eq.detach().numpy() if len(model.eq) else None,
                  b_eq=model.b_eq.detach().numpy() if len(model.eq) else None,
                  bounds=[(None, None)] * model.nvar, method="highs")
    if res.status == 2:
        raise Infeasible("LP infeasible")
    if res.status != 0:
        raise RuntimeError(f"HiGHS status {res.status}: {res.message}")
    x = torch.from_numpy(res.x)
    value = cost @ x
    if len(model.ub):
        y = torch.from_numpy(np.asarray(res.ineqlin.marginals))
        value = value + y @ (model.b_ub - model.A_ub @ x)
    if len(model.eq):
        y = torch.from_numpy(np.asarray(res.eqlin.marginals))
        value = value + y @ (model.b_eq - model.A_eq @ x)
    return res.x, value


def vertex_solution(model: Model, x_np: np.ndarray, quad: Optional[tuple] = None) -> Tensor:
    """Differentiable re-solve on the optimal active set.

    LP  : x = pinv(C) r               with C = [A_eq; A_ub[active]]
    QP  : min 1/2 x^T P x + q^T x on C x = r  via the (pseudo-inverted) KKT matrix, quad=(P, q).
    Coordinates that the active system determines uniquely are smooth functions of the data there.
    """
    x = torch.from_numpy(np.asarray(x_np, float))
    rows, rhs = [model.A_eq], [model.b_eq]
    if len(model.ub):
        slack = (model.b_ub - model.A_ub @ x).detach()
        act = slack.abs() <= ACTIVE_TOL * (1.0 + model.b_ub.detach().abs())
        rows.append(model.A_ub[act])
        rhs.append(model.b_ub[act])
    C, r = torch.cat(rows), torch.cat(rhs)
    if quad is None:
        return torch.linalg.pinv(C) @ r
    P, q = quad
    k = C.shape[0]
    kkt = torch.cat([torch.cat([P, C.T], dim=1), torch.cat([C, torch.zeros(k, k, dtype=F64)], dim=1)])
    sol = torch.linalg.pinv(kkt, hermitian=True) @ torch.cat([-q, r])
    return sol[:model.nvar]


# ------------------------------------------------------------------------------------------------
# the programs
# ------------------------------------------------------------------------------------------------
@dataclass
class SafeSets:
    """Per-environment parameter bundle (one env).  Any of the two blocks may be None."""
    m: int
    n: int
    c_a: Optional[Tensor] = None      # (m,)
    G_a: Optional[Tensor] = None      # (m, p_a)
    c_s: Optional[Tensor] = None      # (n,)
    G_s: Optional[Tensor] = None      # (n, p_s)
    c_f: Optional[Tensor] = None      # (n,)    f(s, a_lin, w_lin)
    B: Optional[Tensor] = None        # (n, m)  df/da
    G_w: Optional[Tensor] = None      # (n, n)  E0 @ G_noise   (constant, baked at wrapper init: Q4)
    a_lin: Optional[Tensor] = None    # (m,)    action_set.center == 0


def _constrain_action(model: Model, a_terms, a_const, sets: SafeSets):
    """feasibility + [SafeAction] point containment + [SafeState] next-state containment for the
    affine action expression  a_k = a_const[k] + sum(a_terms[k]).  (safeguard.py:125-197)"""
    m, n = sets.m, sets.n
    for k in range(m):                                   # -1 <= a <= 1
        model.add_ub(list(a_terms[k]), 1.0 - a_const[k])
        model.add_ub([(i, -c) for i, c in a_terms[k]], 1.0 + a_const[k])
    if sets.G_a is not None:
        add_point_containment(model, "act", a_terms, a_const, sets.c_a, sets.G_a)
    if sets.G_s is not None:
        a_lin = sets.a_lin if sets.a_lin is not None else torch.zeros(m, dtype=F64)
        c1_terms, c1_const = [], []
        for i in range(n):                               # c_f + B (a - a_lin)
            terms = []
            const = sets.c_f[i]
            for k in range(m):
                terms += [(idx, sets.B[i, k] * coef) for idx, coef in a_terms[k]]
                const = const + sets.B[i, k] * (a_const[k] - a_lin[k])
            c1_terms.append(terms)
            c1_const.append(const)
        cols = [([[] for _ in range(n)], [sets.G_w[i, j] for i in range(n)]) for j in range(sets.G_w.shape[1])]
        add_zonotope_containment(model, "st", c1_terms, c1_const, cols, sets.c_s, sets.G_s)


def p1_projection(action: Tensor, sets: SafeSets) -> Tensor:
    """P1, boundary_projection.py:33-57:  argmin ||a - a_s||^2 over the safe set.  Returns a_s (m,).

    m == 1: the feasible a_s form an interval [lo, hi] (two HiGHS LPs), a_s = clamp(a, lo, hi);
    m >= 2: dense active-set on the lifted QP (see ``_qp_active_set``).
    """
    m = sets.m
    model = Model()
    av = model.var("a_s", m)
    zero = torch.zeros((), dtype=F64)
    _constrain_action(model, [[(av.start + k, 1.0)] for k in range(m)], [zero] * m, sets)
    model.finalize()
    if m == 1:
        cost = torch.zeros(model.nvar, dtype=F64)
        cost[av.start] = 1.0
        _, lo = solve_lp(model, cost)
        _, neg_hi = solve_lp(model, -cost)
        hi = -neg_hi
        return torch.minimum(torch.maximum(action, lo.reshape(1)), hi.reshape(1)), lo, hi
    P = torch.zeros(model.nvar, model.nvar, dtype=F64)
    q = torch.zeros(model.nvar, dtype=F64)
    for k in range(m):
        P[av.start + k, av.start + k] = 1.0
    q = q + torch.cat([-action, torch.zeros(model.nvar - m, dtype=F64)])
    x_np = _qp_active_set(model, P.numpy(), q.detach().numpy())
    x = vertex_solution(model, x_np, quad=(P, q))
    return x[av], None, None


def p3_ray_distance(direction: Tensor, center: Tensor, gen: Tensor) -> Tensor:
    """P3, ray_mask.py:215-240:  max t >= 0  s.t.  center + t d in Z(center, gen).  Returns t ()."""
    m = direction.shape[0]
    model = Model()
    tv = model.var("t", 1)
    model.add_ub([(tv.start, -1.0)], 0.0)
    zero = torch.zeros((), dtype=F64)
    add_point_containment(model, "z", [[(tv.start, direction[k])] for k in range(m)], [center[k] + zero for k in range(m)],
                          center, gen)
    model.finalize()
    cost = torch.zeros(model.nvar, dtype=F64)
    cost[tv.start] = -1.0
    _, neg_t = solve_lp(model, cost)
    return -neg_t


def p4_implicit_distance(start: Tensor, direction: Tensor, sets: SafeSets) -> Tensor:
    """P4, ray_mask.py:275-334:  max t >= 0  s.t.  start + t d  feasible, action-safe and state-safe."""
    m = sets.m
    model = Model()
    tv = model.var("t", 1)
    model.add_ub([(tv.start, -1.0)], 0.0)
    _constrain_action(model, [[(tv.start, direction[k])] for k in range(m)], [start[k] for k in range(m)], sets)
    model.finalize()
    cost = torch.zeros(model.nvar, dtype=F64)
    cost[tv.start] = -1.0
    _, neg_t = solve_lp(model, cost)
    return -neg_t


def p2_expansion(directions: Tensor, sets: SafeSets, B0: Optional[Tensor]):
    """P2, ray_mask.py:145-191: max geo_mean(len) over centre c and lengths len >= 0 of Z(c, D diag(len)).

    Solved in two stages: scipy SLSQP on  max sum(log len)  for the active set, then a torch Newton
    polish of the KKT system (gives machine precision and the implicit derivative).  Returns (c, len).
    """
    m = sets.m
    q = directions.shape[1]
    model = Model()
    cv = model.var("c", m)
    lv = model.var("len", q)
    absD = directions.abs()
    for k in range(m):     # -1 <= c -+ |G| 1 <= 1   (ray_mask.py:157-158); |G_kj| = |D_kj| len_j as len >= 0
        model.add_ub([(cv.start + k, 1.0)] + [(lv.start + j, absD[k, j]) for j in range(q)], 1.0)
        model.add_ub([(cv.start + k, -1.0)] + [(lv.start + j, absD[k, j]) for j in range(q)], 1.0)
    for j in range(q):
        model.add_ub([(lv.start + j, -1.0)], 0.0)
    zero = torch.zeros((), dtype=F64)
    c_terms = [[(cv.start + k, 1.0)] for k in range(m)]
    if sets.G_a is not None:      # enc(Z(c, G) in Z(c_a, G_a))
        cols = [([[(lv.start + j, directions[k, j])] for k in range(m)], [zero] * m) for j in range(q)]
        add_zonotope_containment(model, "act", c_terms, [zero] * m, cols, sets.c_a, sets.G_a)
    if sets.G_s is not None:      # enc(Z(c_f + B c, [B0 G, G_w]) in Z(c_s, G_s))   (ray_mask.py:165-179)
        n = sets.n
        c1_terms = [[(cv.start + k, sets.B[i, k]) for k in range(m)] for i in range(n)]
        c1_const = [sets.c_f[i] for i in range(n)]
        cols = []
        B0D = B0 @ directions                       # (n, q); column j scales with len_j
        for j in range(q):
            cols.append(([[(lv.start + j, B0D[i, j])] for i in range(n)], [zero] * n))
        for j in range(sets.G_w.shape[1]):
            cols.append(([[] for _ in range(n)], [sets.G_w[i, j] for i in range(n)]))
        add_zonotope_containment(model, "st", c1_terms, c1_const, cols, sets.c_s, sets.G_s)
    model.finalize()
    x_np = _max_sum_log(model, lv)
    x = _polish_sum_log(model, x_np, lv)
    return x[cv], x[lv]


# ------------------------------------------------------------------------------------------------
# dense helpers for the non-LP programs
# ------------------------------------------------------------------------------------------------
def _phase1(model: Model) -> np.ndarray:
    """A strictly feasible-ish starting point: maximise the minimum inequality slack."""
    nv = model.nvar
    A_ub, b_ub = model.A_ub.detach().numpy(), model.b_ub.detach().numpy()
    A_eq, b_eq = model.A_eq.detach().numpy(), model.b_eq.detach().numpy()
    cost = np.zeros(nv + 1); cost[-1] = -1.0
    res = linprog(cost, A_ub=np.hstack([A_ub, np.ones((A_ub.shape[0], 1))]), b_ub=b_ub,
                  A_eq=np.hstack([A_eq, np.zeros((A_eq.shape[0], 1))]) if A_eq.shape[0] else None,
                  b_eq=b_eq if A_eq.shape[0] else None,
                  bounds=[(None, None)] * nv + [(None, 1.0)], method="highs")
    if res.status != 0 or res.x[-1] < -1e-9:
        raise Infeasible("no feasible point")
    return res.x[:nv]


def _qp_active_set(model: Model, P: np.ndarray, q: np.ndarray) -> np.ndarray:
    """min 1/2 x^T P x + q^T x (P PSD, possibly singular) via scipy SLSQP from a phase-1 point, then
    the caller polishes on the identified active set."""
    x0 = _phase1(model)
    A_ub, b_ub = model.A_ub.detach().numpy(), model.b_ub.detach().numpy()
    A_eq, b_eq = model.A_eq.detach().numpy(), model.b_eq.detach().numpy()
    cons = [{"type": "ineq", "fun": lambda x: b_ub - A_ub @ x, "jac": lambda x: -A_ub}]
    if A_eq.shape[0]:
        cons.append({"type": "eq", "fun": lambda x: A_eq @ x - b_eq, "jac": lambda x: A_eq})
    res = minimize(lambda x: 0.5 * x @ P @ x + q @ x, x0, jac=lambda x: P @ x + q, constraints=cons,
                   method="SLSQP", options={"maxiter": 500, "ftol": 1e-15})
    x = res.x
    # snap: constraints within 1e-7 are declared active, then exact re-solve happens in torch
    slack = b_ub - A_ub @ x
    act = slack <= 1e-7 * (1.0 + np.abs(b_ub))
    C = np.vstack([A_eq, A_ub[act]])
    r = np.concatenate([b_eq, b_ub[act]])
    k = C.shape[0]
    kkt = np.block([[P, C.T], [C, np.zeros((k, k))]])
    sol = np.linalg.pinv(kkt, hermitian=True) @ np.concatenate([-q, r])
    x_new = sol[:model.nvar]
    # project back exactly onto the active rows so vertex_solution() re-identifies the same set
    return x_new


def _max_sum_log(model: Model, lv: slice) -> np.ndarray:
    x0 = _phase1(model)
    A_ub, b_ub = model.A_ub.detach().numpy(), model.b_ub.detach().numpy()
    A_eq, b_eq = model.A_eq.detach().numpy(), model.b_eq.detach().numpy()
    cons = [{"type": "ineq", "fun": lambda x: b_ub - A_ub @ x, "jac": lambda x: -A_ub}]
    if A_eq.shape[0]:
        cons.append({"type": "eq", "fun": lambda x: A_eq @ x - b_eq, "jac": lambda x: A_eq})

    def obj(x):
        return -np.sum(np.log(np.maximum(x[lv], 1e-300)))

    def grad(x):
        g = np.zeros_like(x)
        g[lv] = -1.0 / np.maximum(x[lv], 1e-300)
        return g
    res = minimize(obj, x0, jac=grad, constraints=cons, method="SLSQP", options={"maxiter": 1000, "ftol": 1e-16})
    return res.x


def _polish_sum_log(model: Model, x_np: np.ndarray, lv: slice, iters: int = 30) -> Tensor:
    """Newton on the KKT system of  max sum(log len)  restricted to the active rows; the last step is
    taken with the differentiable data so autograd sees the implicit derivative."""
    x = torch.from_numpy(np.asarray(x_np, float))
    slack = (model.b_ub - model.A_ub @ x).detach()
    act = slack <= 1e-6 * (1.0 + model.b_ub.detach().abs())
    C = torch.cat([model.A_eq, model.A_ub[act]])
    r = torch.cat([model.b_eq, model.b_ub[act]])
    k = C.shape[0]
    nv = model.nvar

    def newton(x_cur: Tensor, C_: Tensor, r_: Tensor) -> Tensor:
        g = torch.zeros(nv, dtype=F64)
        h = torch.zeros(nv, dtype=F64)
        g[lv] = -1.0 / x_cur[lv]
        h[lv] = 1.0 / x_cur[lv] ** 2
        kkt = torch.cat([torch.cat([torch.diag(h), C_.T], dim=1),
                         torch.cat([C_, torch.zeros(k, k, dtype=F64)], dim=1)])
        rhs = torch.cat([-g + h * x_cur, r_])
        return (torch.linalg.pinv(kkt, hermitian=True) @ rhs)[:nv]

    with torch.no_grad():
        for _ in range(iters):
            x = newton(x, C, r)
    return newton(x.detach(), C, r)


# ------------------------------------------------------------------------------------------------
# parametrised fast path (what cvxpylayers does: canonicalise once, refill the data per environment)
# ------------------------------------------------------------------------------------------------
class ParametricP1M1:
    """P1 for m = 1 with the lifted LP canonicalised ONCE: A(theta) = A0 + sum_i theta_i A_i with
    theta = (c_f, B[:, 0]) -- exactly affine, so the decomposition is read off n_theta + 1 builds of the
    generic model above.  Per environment: two HiGHS solves; gradients by the Lagrangian identity."""

    def __init__(self, template: SafeSets):
        self.n = template.n
        self.state = template.G_s is not None
        n_theta = 2 * self.n if self.state else 0
        self.n_theta = n_theta

        def build(theta: Tensor):
            s = SafeSets(m=1, n=template.n, c_a=template.c_a, G_a=template.G_a, c_s=template.c_s, G_s=template.G_s,
                         G_w=template.G_w, a_lin=template.a_lin)
            if self.state:
                s.c_f, s.B = theta[:self.n], theta[self.n:].reshape(self.n, 1)
            model = Model()
            av = model.var("a_s", 1)
            _constrain_action(model, [[(av.start, 1.0)]], [torch.zeros((), dtype=F64)], s)
            model.finalize()
            return model, av.start

        base, self.a_idx = build(torch.zeros(n_theta, dtype=F64))
        self.nvar = base.nvar
        self.A_ub0, self.b_ub0 = base.A_ub.numpy().copy(), base.b_ub.numpy().copy()
        self.A_eq0, self.b_eq0 = base.A_eq.numpy().copy(), base.b_eq.numpy().copy()
        self.dA_ub, self.db_ub, self.dA_eq, self.db_eq = [], [], [], []
        for i in range(n_theta):
            e = torch.zeros(n_theta, dtype=F64)
            e[i] = 1.0
            m_i, _ = build(e)
            self.dA_ub.append(m_i.A_ub.numpy() - self.A_ub0)
            self.db_ub.append(m_i.b_ub.numpy() - self.b_ub0)
            self.dA_eq.append(m_i.A_eq.numpy() - self.A_eq0)
            self.db_eq.append(m_i.b_eq.numpy() - self.b_eq0)
        if n_theta:
            self.dA_ub, self.db_ub = np.stack(self.dA_ub), np.stack(self.db_ub)
            self.dA_eq, self.db_eq = np.stack(self.dA_eq), np.stack(self.db_eq)
        self.cost = np.zeros(self.nvar)
        self.cost[self.a_idx] = 1.0

    def bounds(self, theta: np.ndarray):
        """(lo, hi, dlo/dtheta, dhi/dtheta) for one environment; raises Infeasible."""
        if self.n_theta:
            A_ub = self.A_ub0 + np.tensordot(theta, self.dA_ub, axes=1)
            b_ub = self.b_ub0 + theta @ self.db_ub
            A_eq = self.A_eq0 + np.tensordot(theta, self.dA_eq, axes=1)
            b_eq = self.b_eq0 + theta @ self.db_eq
        else:
            A_ub, b_ub, A_eq, b_eq = self.A_ub0, self.b_ub0, self.A_eq0, self.b_eq0
        out = []
        for sign in (1.0, -1.0):
            res = linprog(sign * self.cost, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq if len(b_eq) else None,
                          b_eq=b_eq if len(b_eq) else None, bounds=[(None, None)] * self.nvar, method="highs")
            if res.status == 2:
                raise Infeasible("LP infeasible")
            if res.status != 0:
                raise RuntimeError(f"HiGHS status {res.status}: {res.message}")
            grad = np.zeros(self.n_theta)
            if self.n_theta:
                y_ub, y_eq = np.asarray(res.ineqlin.marginals), np.asarray(res.eqlin.marginals)
                grad = (self.db_ub - self.dA_ub @ res.x) @ y_ub + (self.db_eq - self.dA_eq @ res.x) @ y_eq
            out.append((sign * res.fun, sign * grad))
        return out[0][0], out[1][0], out[0][1], out[1][1]


class _IntervalFn(torch.autograd.Function):
    """(lo, hi) of every environment as a differentiable function of theta (N, n_theta)."""

    @staticmethod
    def forward(ctx, theta: Tensor, prog: ParametricP1M1):
        th = theta.detach().numpy()
        N = th.shape[0]
        lo, hi = np.full(N, np.nan), np.full(N, np.nan)
        glo, ghi = np.zeros_like(th), np.zeros_like(th)
        infeasible = np.zeros(N, dtype=bool)
        for i in range(N):
            if not np.isfinite(th[i]).all():       # a diverged environment: the reference's solver raises on non-finite
                infeasible[i] = True               # parameters; under strict=False the row is reported infeasible (NaN action)
                continue
            try:
                lo[i], hi[i], glo[i], ghi[i] = prog.bounds(th[i])
            except Infeasible:
                infeasible[i] = True
        ctx.save_for_backward(torch.from_numpy(glo), torch.from_numpy(ghi))
        ctx.infeasible = infeasible
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(infeasible)

    @staticmethod
    def backward(ctx, g_lo, g_hi, _g_inf):
        glo, ghi = ctx.saved_tensors
        return g_lo.unsqueeze(1) * glo + g_hi.unsqueeze(1) * ghi, None


def p1_projection_batch_m1(action: Tensor, prog: ParametricP1M1, theta: Tensor):
    """Batched P1 (m = 1): returns (a_s (N,1) with NaN rows where infeasible, infeasible mask)."""
    lo, hi, infeasible = _IntervalFn.apply(theta, prog)
    a_s = torch.minimum(torch.maximum(action[:, 0], lo), hi)
    a_s = torch.where(infeasible, torch.full_like(a_s, float("nan")), a_s)
    return a_s.unsqueeze(1), infeasible



def p5_seeker_generator(state: Tensor, obstacles, box_min: Tensor, box_max: Tensor) -> Tensor:
    """P5, navigate_seeker.py:415-477 for ONE environment: generator U diag(len) of the safe action zonotope,
    len = argmax geo_mean(len) under state-box feasibility, len <= 1 and one support constraint per obstacle.
    Solved generically (SLSQP on sum(log len), then an exact re-solve on the identified active set)."""
    U = torch.eye(2, dtype=F64)
    min_dist = torch.linalg.vector_norm(state.abs() - box_max)
    rows = []
    for center, radius in obstacles:
        direction = center - state
        distance = torch.linalg.vector_norm(direction)
        to_obs = direction / distance
        distance = distance - radius
        if distance < min_dist:
            U = torch.stack([to_obs, torch.stack([to_obs[1], -to_obs[0]])], dim=1)
            min_dist = distance
    A, b = [], []
    for d in range(2):
        A.append(U[d].abs()); b.append(state[d] - box_min[d])
        A.append(U[d].abs()); b.append(box_max[d] - state[d])
    A.append(torch.tensor([1.0, 0.0], dtype=F64)); b.append(torch.tensor(1.0, dtype=F64))
    A.append(torch.tensor([0.0, 1.0], dtype=F64)); b.append(torch.tensor(1.0, dtype=F64))
    for center, radius in obstacles:
        direction = center - state
        distance = torch.linalg.vector_norm(direction)
        A.append((direction / distance @ U).abs()); b.append(distance - radius)
    A, b = torch.stack(A).numpy(), torch.stack(b).numpy()
    x0 = np.full(2, 0.25 * min(1.0, float(np.min(b / np.maximum(A.sum(1), 1e-12)))))
    res = minimize(lambda x: -np.sum(np.log(x)), x0, jac=lambda x: -1.0 / x,
                   constraints=[{"type": "ineq", "fun": lambda x: b - A @ x, "jac": lambda x: -A}],
                   bounds=[(1e-12, None)] * 2, method="SLSQP", options={"maxiter": 500, "ftol": 1e-16})
    x = res.x
    act = np.where(b - A @ x <= 1e-7 * (1 + np.abs(b)))[0]
    if len(act) >= 2:                      # vertex
        x = np.linalg.lstsq(A[act], b[act], rcond=None)[0]
    elif len(act) == 1:                    # edge optimum of the product: x_j = b / (2 a_j)
        x = b[act[0]] / (2 * A[act[0]])
    return U * torch.from_numpy(x)
