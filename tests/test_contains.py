"""Containment verdicts per environment on the device (sgbpo_contains; reference programs P9, zonotope.py:143-239).

CPU: the facet enumeration (sets/verdicts.py) reproduces the oracle's LP value for points (gauge duality) and the reference's
own hand-written contained / not-contained cases.  GPU: the kernel's value equals the oracle LP value for points (1e-9, f64), its
verdict equals the oracle's LP verdict for zonotopes with shared inner generators (Sadraddini-Tedrake encoding), and with
per-environment inner generators it equals the LP value for parallelotope outer sets and is a lower bound of it otherwise."""
import numpy as np
import pytest
import torch

from oracle import sets as osets


def _rand_zonotope(rs, n, p):
    return rs.normal(size=n), rs.normal(size=(n, p)) * rs.uniform(0.3, 1.5, size=p)


@pytest.mark.parametrize("n,p", [(1, 2), (2, 2), (2, 5), (3, 3), (3, 6), (4, 4), (4, 7)])
def test_facets_give_the_point_lp_value(n, p):
    from safegbpo_b200.sets.verdicts import zonotope_facets
    rs = np.random.default_rng(10 * n + p)
    c, G = _rand_zonotope(rs, n, p)
    H = zonotope_facets(G)
    for _ in range(12):
        x = c + G @ rs.uniform(-1.4, 1.4, size=p)
        want = osets.zonotope_point_norm(c, G, x)
        got = float(np.max(H @ (x - c)))
        assert abs(got - want) <= 1e-8 * max(1.0, want), (got, want)


def test_reference_hand_cases():
    """sets_test.py: points of the unit square zonotope [[1, 0], [0, 1]] and of the rotated one."""
    from safegbpo_b200.sets.verdicts import zonotope_facets
    H = zonotope_facets(np.eye(2))
    assert np.max(H @ np.array([0.5, -1.0])) <= 1.0 and np.max(H @ np.array([1.0, 1.0])) <= 1.0
    assert np.max(H @ np.array([1.0001, 0.0])) > 1.0
    H = zonotope_facets(np.array([[1.0, 1.0], [1.0, -1.0]]))          # diamond with vertices (+-2, 0), (0, +-2)
    assert np.max(H @ np.array([2.0, 0.0])) <= 1.0 + 1e-12 and np.max(H @ np.array([1.5, 1.0])) > 1.0


def test_shared_generator_facets_give_the_zonotope_lp_verdict():
    """Zonotope program with fixed inner generators: the facets of the LP's feasible region in the centre offset (what DeviceVerdict
    uploads) decide exactly like the reference's Sadraddini-Tedrake LP (zonotope.py:213-236)."""
    from safegbpo_b200.consts import merge_parallel_generators, polytope_facets
    rs = np.random.default_rng(3)
    for n, p, q in [(2, 4, 2), (3, 5, 3)]:
        c, G = _rand_zonotope(rs, n, p)
        Gi = 0.15 * rs.normal(size=(n, q))
        centers = c + (G @ rs.uniform(-1.25, 1.25, size=(p, 60))).T
        H = polytope_facets(merge_parallel_generators(G), Gi)
        value = np.max((centers - c) @ H.T, axis=1)
        want = np.array([osets.zonotope_in_zonotope_norm(ci, Gi, c, G) for ci in centers])
        clear = np.abs(want - 1.0) > 1e-7
        assert np.array_equal((value <= 1.0)[clear], (want <= 1.0)[clear]), (n, p, q)
        assert 0.05 < (want <= 1.0).mean() < 0.99


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n,p", [(2, 4), (4, 6), (6, 7)])
def test_point_values_vs_oracle_lp(dev, dtype, n, p):
    from safegbpo_b200.sets.verdicts import DeviceVerdict
    rs = np.random.default_rng(n * 100 + p)
    c, G = _rand_zonotope(rs, n, p)
    N = 300
    pts = c + (G @ rs.uniform(-1.3, 1.3, size=(p, N))).T
    v = DeviceVerdict(torch.tensor(c, dtype=dtype, device=dev), torch.tensor(G, dtype=dtype, device=dev))
    value, inside = v(torch.tensor(pts, dtype=dtype, device=dev))
    want = np.array([osets.zonotope_point_norm(c, G, x) for x in pts])
    tol = 1e-9 if dtype == torch.float64 else 2e-5
    np.testing.assert_allclose(value.cpu().numpy(), want, rtol=tol, atol=tol)
    clear = np.abs(want - 1.0) > 10 * tol
    assert np.array_equal(inside.cpu().numpy()[clear], (want <= 1.0)[clear])
    assert 0.2 < (want <= 1.0).mean() < 0.95                            # both verdicts occur


@pytest.mark.gpu
def test_zonotope_verdicts_with_shared_generators_vs_oracle_lp(dev):
    """The constant-safe-set case: every environment's inner zonotope has the same generators (e.g. the noise set around the
    next nominal state).  Verdict = the reference's Sadraddini-Tedrake LP verdict."""
    from safegbpo_b200.sets.verdicts import DeviceVerdict
    rs = np.random.default_rng(3)
    for n, p, q in [(2, 4, 2), (3, 5, 3), (4, 6, 2)]:
        c, G = _rand_zonotope(rs, n, p)
        Gi = 0.15 * rs.normal(size=(n, q))
        N = 120
        centers = c + (G @ rs.uniform(-1.25, 1.25, size=(p, N))).T
        v = DeviceVerdict(torch.tensor(c, device=dev), torch.tensor(G, device=dev), torch.tensor(Gi, device=dev))
        value, inside = v(torch.tensor(centers, device=dev))
        want = np.array([osets.zonotope_in_zonotope_norm(ci, Gi, c, G) for ci in centers])
        clear = np.abs(want - 1.0) > 1e-7
        assert np.array_equal(inside.cpu().numpy()[clear], (want <= 1.0)[clear]), (n, p, q)
        assert 0.1 < (want <= 1.0).mean() < 0.99                       # both verdicts occur
        on = np.abs(value.cpu().numpy() - 1.0) < 1e-3                   # on the boundary the two values coincide
        np.testing.assert_allclose(want[on], 1.0, atol=5e-3)


@pytest.mark.gpu
def test_per_environment_generators(dev):
    from safegbpo_b200.sets.verdicts import DeviceVerdict
    rs = np.random.default_rng(4)
    n, q, N = 4, 5, 150
    # parallelotope outer set: the exact value IS the LP value
    c, G = rs.normal(size=n), rs.normal(size=(n, n)) + 2 * np.eye(n)
    centers = c + (G @ rs.uniform(-0.9, 0.9, size=(n, N))).T
    gens = 0.12 * rs.normal(size=(N, n, q))
    v = DeviceVerdict(torch.tensor(c, device=dev), torch.tensor(G, device=dev))
    value, inside = v(torch.tensor(centers, device=dev), torch.tensor(gens, device=dev))
    want = np.array([osets.zonotope_in_zonotope_norm(centers[i], gens[i], c, G) for i in range(N)])
    np.testing.assert_allclose(value.cpu().numpy(), want, rtol=1e-9, atol=1e-9)
    assert np.array_equal(inside.cpu().numpy(), want <= 1.0)
    # general outer zonotope: exact containment, never more conservative than the reference's sufficient LP
    c, G = _rand_zonotope(rs, 3, 6)
    centers = c + (G @ rs.uniform(-0.8, 0.8, size=(6, N))).T
    gens = 0.1 * rs.normal(size=(N, 3, 4))
    v = DeviceVerdict(torch.tensor(c, device=dev), torch.tensor(G, device=dev))
    value, inside = v(torch.tensor(centers, device=dev), torch.tensor(gens, device=dev))
    want = np.array([osets.zonotope_in_zonotope_norm(centers[i], gens[i], c, G) for i in range(N)])
    assert (value.cpu().numpy() <= want + 1e-9).all()
    assert not (~inside.cpu().numpy() & (want <= 1.0 - 1e-9)).any()     # LP says contained => the exact test agrees
    # exactness: every vertex direction's support point is inside the outer zonotope iff value <= 1
    i = int(np.argmax(value.cpu().numpy()))
    worst = max(osets.zonotope_point_norm(c, G, centers[i] + gens[i] @ (2.0 * np.array(bits) - 1.0)) for bits in np.ndindex(2, 2, 2, 2))
    assert abs(worst - float(value[i])) < 1e-8


@pytest.mark.gpu
def test_per_environment_centres_and_errors(dev):
    import ctypes as C
    from safegbpo_b200 import _lib
    from safegbpo_b200.sets.verdicts import DeviceVerdict
    G = torch.tensor([[1.0, 0.5], [0.0, 1.0]], device=dev, dtype=torch.float64)
    centres = torch.tensor([[0.0, 0.0], [10.0, 0.0]], device=dev, dtype=torch.float64)
    v = DeviceVerdict(centres, G)
    value, inside = v(torch.tensor([[1.4, 0.9], [1.4, 0.9]], device=dev, dtype=torch.float64))
    assert inside.tolist() == [True, False]
    nan_value, nan_inside = v(torch.tensor([[float("nan"), 0.0], [10.0, 0.0]], device=dev, dtype=torch.float64))
    assert nan_inside.tolist() == [False, True]
    a = _lib.ContainsArgs()
    assert _lib.lib().sgbpo_contains(1, C.byref(a), None) == -1 and b"sgbpo_contains" in _lib.lib().sgbpo_last_error()
