"""CPU: pin the oracle's convex programs with what the reference itself offers at the solver boundary
(no numeric safe action or gradient is pinned by the reference's tests -- 'parity unpinned', see
oracle/__init__.py): zonotope verdict vectors (tests/sets_test.py), the paper's known answers
(paper_archive/equation_check.ipynb), and independent cross-checks (closed forms, finite differences)."""
import numpy as np
import pytest
import torch

from oracle import programs as P
from oracle.sets import zonotope_contains_point, zonotope_contains_zonotope

F64 = torch.float64


def test_zonotope_verdicts_from_reference_sets_test():
    """tests/sets_test.py:96-101, 365-403: outer Z c=(1,2), G=[[1,1.3,2.4,-1],[1,-1,0,1.5]]."""
    c2, g2 = [1.0, 2.0], [[1.0, 1.3, 2.4, -1.0], [1.0, -1.0, 0.0, 1.5]]
    inner = 0.9 * np.array([[1.0, 0.5, 0.2], [0.5, 1.0, 0.3]])
    assert zonotope_contains_zonotope([1.0, 2.0], inner, c2, g2)
    assert zonotope_contains_zonotope([3.0, 1.0], inner, c2, g2)
    assert not zonotope_contains_zonotope([-3.0, 3.0], inner / 0.9, c2, g2)
    assert not zonotope_contains_zonotope([4.5, 0.5], inner / 0.9, c2, g2)
    # every sampled point of a zonotope is contained (sets_test.py:79-101 idea)
    rng = np.random.default_rng(0)
    G = np.array(g2)
    for _ in range(50):
        p = np.array(c2) + G @ rng.uniform(-1, 1, 4)
        assert zonotope_contains_point(c2, G, p)
    assert not zonotope_contains_point(c2, G, [10.0, 2.0])


def _action_only(m, G, c=None):
    return P.SafeSets(m=m, n=0, c_a=torch.zeros(m, dtype=F64) if c is None else c, G_a=G)


def _p1(a, sets):
    return P.p1_projection(a, sets)[0]


def test_eq13_14_boundary_ray():
    """equation_check.ipynb cell 7: G = I2, actions on the ray (1,0) + t (1,0) project to (1,0) and the
    directional derivative along the ray is 0."""
    sets = _action_only(2, torch.eye(2, dtype=F64))
    for t in np.linspace(0.05, 1, 6):
        a = torch.tensor([1.0 + t, 0.0], dtype=F64, requires_grad=True)
        a_s = _p1(a, sets)
        np.testing.assert_allclose(a_s.detach().numpy(), [1.0, 0.0], atol=1e-9)
        jac = torch.stack([torch.autograd.grad(a_s[i], a, retain_graph=True)[0] for i in range(2)])
        assert abs(float(torch.ones(2, dtype=F64) @ jac @ torch.tensor([1.0, 0.0], dtype=F64))) < 1e-8


def test_eq71_projector_onto_inactive_generators():
    """cells 24-25: J = G_I (G_I^T G_I)^+ G_I^T; example a = (2,0), G = I2 -> J = diag(0,1)."""
    sets = _action_only(2, torch.eye(2, dtype=F64))
    a = torch.tensor([2.0, 0.0], dtype=F64, requires_grad=True)
    a_s = _p1(a, sets)
    jac = torch.stack([torch.autograd.grad(a_s[i], a, retain_graph=True)[0] for i in range(2)])
    np.testing.assert_allclose(jac.numpy(), np.diag([0.0, 1.0]), atol=1e-8)


def test_eq15_rank_full_iff_safe():
    """cell 9: rank(J) = d iff the action is already safe (d = 2, G ~ U(0, .5)^{2x4})."""
    gen = torch.Generator().manual_seed(0)
    G = torch.rand(2, 4, generator=gen, dtype=F64) * 0.5
    sets = _action_only(2, G)
    for _ in range(6):
        a = torch.rand(2, generator=gen, dtype=F64, requires_grad=True)
        a_s = _p1(a, sets)
        jac = torch.stack([torch.autograd.grad(a_s[i], a, retain_graph=True)[0] for i in range(2)])
        safe = bool(torch.allclose(a_s.detach(), a.detach(), atol=1e-7))
        rank = np.linalg.matrix_rank(jac.numpy(), tol=1e-6)
        assert (rank == 2) == safe


def test_p3_is_reciprocal_zonotope_norm():
    """P3: t* = sum|g_j| for m = 1 (pendulum ctrl zonotope -> 1.000002941) and 1/||G^-1 d||_inf
    for an invertible 2-D generator."""
    from oracle.dynamics import load_assets
    g = load_assets()["pendulum_ctrl_generators"]
    t = P.p3_ray_distance(torch.ones(1, dtype=F64), torch.zeros(1, dtype=F64), g)
    np.testing.assert_allclose(float(t), float(g.abs().sum()), rtol=1e-12)
    G = torch.tensor([[2.0, 0.5], [0.0, 1.5]], dtype=F64)
    d = torch.tensor([0.6, 0.8], dtype=F64)
    t2 = P.p3_ray_distance(d, torch.tensor([0.3, -0.2], dtype=F64), G)
    np.testing.assert_allclose(float(t2), 1.0 / float(torch.linalg.solve(G, d).abs().max()), rtol=1e-10)


def test_p1_gradient_matches_finite_differences():
    from oracle.dynamics import OracleEnv
    env = OracleEnv("BalanceCartPole", 1, 10)
    env.state = torch.tensor([[2.0, 0.4, 9.95, -3.0]], dtype=F64)
    _, _, _, e_mat = env.linear_dynamics()

    def hi_of(state):
        env.state = state
        c, _, B, _ = env.linear_dynamics()
        sets = P.SafeSets(m=1, n=4, c_a=env.ctrl[0], G_a=env.ctrl[1], c_s=env.rci[0], G_s=env.rci[1], c_f=c[0], B=B[0],
                          G_w=torch.zeros(4, 4, dtype=F64))
        return P.p1_projection(torch.tensor([0.9], dtype=F64), sets)[2]

    s = env.state.clone().requires_grad_(True)
    hi = hi_of(s)
    assert float(hi) < 1.0                      # the velocity row binds
    grad = torch.autograd.grad(hi, s)[0][0]
    fd = []
    x0 = s.detach().clone()
    for i in range(4):
        d = torch.zeros(1, 4, dtype=F64); d[0, i] = 1e-6
        fd.append((float(hi_of(x0 + d)) - float(hi_of(x0 - d))) / 2e-6)
    np.testing.assert_allclose(grad.numpy(), fd, rtol=1e-5, atol=1e-7)


def test_parametric_fast_path_equals_generic():
    from common import make_case, oracle_step
    a = make_case("SwingUpCartPole", "boundary_projection", 12, seed=9)
    b = make_case("SwingUpCartPole", "boundary_projection", 12, seed=9)
    b["sg"].fast = False
    ra, rb = oracle_step(a), oracle_step(b)
    assert torch.equal(ra["ok"], rb["ok"])
    ok = ra["ok"]
    for k in ("a_exec", "s_next", "gs", "ga"):
        np.testing.assert_allclose(ra[k][ok].numpy(), rb[k][ok].numpy(), rtol=1e-9, atol=1e-10)
