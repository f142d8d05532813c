"""NavigateQuadrotor: the per-step safe state set (programs P6-P8, reference src/envs/navigate_quadrotor.py:436-779), the
collision operator and the reward -- the device math (compiled for the host, tests/hostmath) and the CUDA kernels against
the oracle restatement (oracle/navquad.py: same formulation, independent generic solvers)."""
import ctypes as C

import numpy as np
import pytest
import torch

import hostmath_util as hm
from oracle import navquad as onq
from oracle.dynamics import f_quadrotor as _f_quadrotor_single
from safegbpo_b200 import _lib
from safegbpo_b200.consts import build_consts

f_quadrotor = torch.vmap(_f_quadrotor_single)          # the oracle's dynamics are written per environment (the reference vmaps them)
ENV = _lib.ENV_IDS["NavigateQuadrotor"]
G_W = np.zeros((6, 6))
G_W[0, 3], G_W[3, 3], G_W[1, 4], G_W[4, 4] = 0.05 ** 2 * 0.1, 0.05 * 0.1, 0.05 ** 2 * 0.1, 0.05 * 0.1     # E_0 G_noise


def _cases(N, seed, k_obs=2):
    g = np.random.default_rng(seed)
    half = onq.BOX_MAX
    s = (g.random((N, 6)) * 2 - 1) * half * np.array([0.98, 0.98, 0.9, 0.95, 0.95, 0.5])
    s[: N // 6, 0] = np.sign(s[: N // 6, 0]) * (7.9 + 0.1 * g.random(N // 6))          # hugging a wall, moving along it
    goal = (g.random((N, 2)) * 2 - 1) * 8
    aux = np.zeros((N, 13))
    aux[:, 0:2] = goal
    obstacles = []
    for i in range(N):
        obs_i = []
        for j in range(k_obs):
            c = s[i, :2] + (g.random(2) * 2 - 1) * 5.0
            r = 0.5 + 1.5 * g.random()
            if np.linalg.norm(c - s[i, :2]) <= r + 0.05:        # the quadrotor is never inside an obstacle at the start of a step
                c = s[i, :2] + (r + 0.3 + g.random()) * np.array([np.cos(j + 1.0), np.sin(j + 1.0)])
            aux[i, 2 + 3 * j: 5 + 3 * j] = [c[0], c[1], r]
            obs_i.append((c, r))
        obstacles.append(obs_i)
    aux[:, 11] = np.linalg.norm(goal - s[:, :2], axis=1) + 0.5
    return s, aux, obstacles


def _consts(k_obs=2, sg=_lib.SG_NONE, gate=2):
    return build_consts("NavigateQuadrotor", sg, gate_mode=gate, dynamic_state_set=True, n_obstacles=k_obs, G_w=G_W)


def _reach(s):
    """(c_f, B) of the oracle: f(s, 0, 0) and df/da by central differences of the oracle dynamics (B is linear in a)."""
    st = torch.from_numpy(s)
    z = torch.zeros(s.shape[0], 2, dtype=torch.float64)
    w = torch.zeros_like(st)
    c_f = f_quadrotor(st, z, w).numpy()
    B = np.zeros((s.shape[0], 6, 2))
    for q in range(2):
        e = z.clone()
        e[:, q] = 1.0
        B[:, :, q] = f_quadrotor(st, e, w).numpy() - c_f
    return c_f, B


def _check_sets(center, gen, status, s, obstacles):
    c_f, B = _reach(s)
    n_ok = 0
    for i in range(s.shape[0]):
        try:
            oc, og, safe_p, safe_v = onq.safe_state_set(c_f[i], B[i], obstacles[i])
        except onq.Infeasible:
            assert status[i] == 1, f"env {i}: the support-point program is infeasible for the oracle"
            continue
        if not (safe_p and safe_v):
            assert status[i] == 1, f"env {i}: the reference falls back to its degenerate blocks here"
            continue
        assert status[i] == 0, f"env {i}"
        # the centre comes from an LP whose optimum may be a whole edge: compare the objective instead when the points differ
        if not np.allclose(center[i], oc, rtol=1e-7, atol=1e-8):
            too_low, too_high = onq.BOX_MIN > c_f[i], onq.BOX_MAX < c_f[i]
            if too_low.any() or too_high.any():
                continue                                        # (shifted centres with a non-unique optimum are skipped below too)
        np.testing.assert_allclose(center[i], oc, rtol=1e-7, atol=1e-8, err_msg=f"centre {i}")
        # generators: positions / velocities are U diag(len): compare the zonotope (G G^T) -- column signs are free
        np.testing.assert_allclose(gen[i] @ gen[i].T, og @ og.T, rtol=2e-6, atol=1e-9, err_msg=f"generator {i}")
        n_ok += 1
    return n_ok


def test_host_math_safe_state_set_vs_oracle():
    s, aux, obstacles = _cases(240, seed=3)
    center, gen, status = hm.navquad_safe_state_set(_consts(), s, aux)
    assert _check_sets(center, gen, status, s, obstacles) > 150


def test_host_math_step_collision_and_reward_vs_oracle():
    """Unsafeguarded step: dynamics, collision operator (last obstacle entered, roll kept), clamp, reward with the one-off
    goal bonus and the collision penalty."""
    N = 300
    s, aux, obstacles = _cases(N, seed=5)
    g = np.random.default_rng(9)
    # aim a third of the quadrotors at their first obstacle, fast, from just outside it
    for i in range(0, N, 3):
        c, r = obstacles[i][0]
        d = (c - s[i, :2]) / np.linalg.norm(c - s[i, :2])
        s[i, :2] = c - d * (r + 0.02)
        s[i, 3:5] = d * np.array([0.79, 0.99])
    for i in range(1, N, 7):                                   # and put some next to their goal
        aux[i, 0:2] = s[i, :2] + s[i, 3:5] * 0.05 + 0.01
    a = g.random((N, 2)) * 2 - 1
    w = np.zeros((N, 6))
    k = _consts()
    out = hm.step(ENV, k, s, a, w, aux=aux)
    free = f_quadrotor(torch.from_numpy(s), torch.from_numpy(a), torch.from_numpy(w)).numpy()
    n_hit = n_goal = 0
    for i in range(N):
        nxt, hit = onq.collision_operator(s[i], free[i], obstacles[i])
        nxt = np.clip(nxt, onq.BOX_MIN, onq.BOX_MAX)
        rew, reached = onq.reward(nxt, aux[i, 0:2], aux[i, 11], False, hit)
        np.testing.assert_allclose(out["s_next"][i], nxt, rtol=1e-10, atol=1e-12, err_msg=f"state {i}")
        np.testing.assert_allclose(out["reward"][i], rew, rtol=1e-10, atol=1e-12, err_msg=f"reward {i}")
        assert bool(out["flags"][i] & _lib.FL_COLLIDED) == hit and bool(out["flags"][i] & _lib.FL_REACHED) == reached
        n_hit += hit
        n_goal += reached
    assert n_hit > 20 and n_goal > 10
    # an environment that already reached its goal is not paid twice
    aux2 = aux.copy()
    aux2[:, 12] = 1.0
    out2 = hm.step(ENV, k, s, a, w, aux=aux2)
    assert not (out2["flags"] & _lib.FL_REACHED).any() and (out2["reward"] <= out["reward"] + 1e-12).all()


def test_host_math_boundary_projection_on_the_dynamic_set():
    """P1 on the per-step set: every projected action keeps the linearised next-state zonotope inside THIS step's safe state
    set -- re-checked by the independent LP on the unmerged containment encoding with the ORACLE's generator."""
    from common import p1_feasibility_budget
    N = 200
    s, aux, obstacles = _cases(N, seed=11)
    g = np.random.default_rng(2)
    a = g.random((N, 2)) * 2 - 1
    k = _consts(sg=_lib.SG_BOUNDARY_PROJECTION, gate=2)
    out = hm.step(ENV, k, s, a, np.zeros((N, 6)), aux=aux)
    c_f, B = _reach(s)
    Wn = G_W[:, np.abs(G_W).sum(0) > 0]
    checked = moved = 0
    for i in range(N):
        fl = int(out["flags"][i])
        try:
            oc, og, safe_p, safe_v = onq.safe_state_set(c_f[i], B[i], obstacles[i])
        except onq.Infeasible:
            continue
        if not (safe_p and safe_v):
            assert fl & _lib.FL_INFEASIBLE
            continue
        budget_any = p1_feasibility_budget(og, Wn, oc, c_f[i], B[i])
        if fl & _lib.FL_INFEASIBLE:
            assert budget_any > 1 - 1e-6, f"env {i}: flagged infeasible, independent LP budget {budget_any}"
            continue
        assert fl & _lib.FL_NEXT_IN_SET
        ae = out["a_exec"][i]
        assert np.all(np.abs(ae) <= 1 + 1e-9)
        # containment of the executed action: budget with the action FIXED (B = 0, offset moved into c_f)
        fixed = p1_feasibility_budget(og, Wn, oc, c_f[i] + B[i] @ ae, np.zeros((6, 2)))
        assert fixed <= 1 + 1e-6, f"env {i}: executed action leaves the safe set (budget {fixed})"
        checked += 1
        moved += not np.allclose(ae, a[i], atol=1e-9)
    assert checked > 80 and moved > 20


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_gpu_kernels_vs_host_math(dtype):
    """The CUDA kernels compute what the host-compiled device math computes (which the tests above tie to the oracle)."""
    from safegbpo_b200 import ops
    dev = torch.device("cuda:0")
    N = 4099
    s, aux, obstacles = _cases(N, seed=21)
    g = np.random.default_rng(4)
    a = g.random((N, 2)) * 2 - 1
    k0, k1 = _consts(), _consts(sg=_lib.SG_BOUNDARY_PROJECTION, gate=2)
    t = lambda x: torch.from_numpy(x).to(device=dev, dtype=dtype)
    ndt = np.float64 if dtype == torch.float64 else np.float32
    center, gen, status = hm.navquad_safe_state_set(k0, s, aux, dtype=ndt)
    c_d, g_d, st_d = torch.empty(N, 6, device=dev, dtype=dtype), torch.empty(N, 6, 6, device=dev, dtype=dtype), torch.empty(N, device=dev, dtype=torch.int32)
    s_d, aux_d = t(s), t(aux)            # (kept alive across the launch)
    _lib.check(_lib.lib().sgbpo_navquad_safe_state_set(_lib.dtype_code(dtype), N, C.byref(k0), _lib.ptr(s_d), _lib.ptr(aux_d),
                                                      _lib.ptr(c_d), _lib.ptr(g_d), _lib.ptr(st_d), _lib.stream_ptr()), "navquad")
    tol = dict(rtol=1e-10, atol=1e-12) if dtype == torch.float64 else dict(rtol=2e-4, atol=2e-5)
    same = st_d.cpu().numpy() == status
    assert same.mean() > (0.9999 if dtype == torch.float64 else 0.995)
    ok = same & (status == 0)
    np.testing.assert_allclose(c_d.cpu().numpy()[ok], center[ok], **tol)
    np.testing.assert_allclose(g_d.cpu().numpy()[ok], gen[ok], **tol)
    ref = hm.step(ENV, k1, s, a, np.zeros((N, 6)), aux=aux, dtype=ndt,
                  cot=dict(g_s_next=np.ones((N, 6)), g_a_exec=np.ones((N, 2)), g_obs=np.ones((N, 8)), g_reward=np.ones(N)))
    sd, ad = t(s).requires_grad_(True), t(a).requires_grad_(True)
    s_next, a_exec, obs, rew, fl = ops.safeguarded_step(sd, ad, torch.zeros(N, 6, device=dev, dtype=dtype), ENV, k1, aux=t(aux))
    flags_equal = fl.cpu().numpy() == ref["flags"]
    assert flags_equal.mean() > (0.9999 if dtype == torch.float64 else 0.99)
    good = flags_equal & ((ref["flags"] & _lib.FL_INFEASIBLE) == 0)
    (s_next.sum(1) + a_exec.sum(1) + obs.sum(1) + rew)[torch.from_numpy(good).to(dev)].sum().backward()
    np.testing.assert_allclose(a_exec.detach().cpu().numpy()[good], ref["a_exec"][good], **tol)
    np.testing.assert_allclose(s_next.detach().cpu().numpy()[good], ref["s_next"][good], **tol)
    np.testing.assert_allclose(rew.detach().cpu().numpy()[good], ref["reward"][good], **tol)
    if dtype == torch.float64:
        np.testing.assert_allclose(sd.grad.cpu().numpy()[good], ref["gs"][good], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(ad.grad.cpu().numpy()[good], ref["ga"][good], rtol=1e-8, atol=1e-9)


@pytest.mark.gpu
def test_env_class_rollout_and_shac_episode():
    """The drop-in class: reset (goal / obstacle rejection sampling), observation layout, a boundary-projection SHAC episode."""
    from safegbpo_b200.envs.navigate_quadrotor import NavigateQuadrotorEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.learning_algorithms.shac import SHAC
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda:0")
    try:
        torch.manual_seed(0)
        env = NavigateQuadrotorEnv(num_envs=64, num_steps=240, num_obstacles=1, min_radius=2.0, max_radius=3.0)
        obs, _ = env.reset(seed=1)
        assert obs.shape == (64, 11) and env.obs_dim == 11
        pos, goal, ball = env.state[:, :2], env.goal, env.obstacles[0]
        assert bool((torch.linalg.vector_norm(goal - pos, dim=1) >= 2 * env.max_radius).all())
        assert not bool(ball.contains(pos).any()) and not bool(ball.contains(goal).any())
        z = env.safe_state_set()
        assert z.center.shape == (64, 6) and z.generator.shape == (64, 6, 6)
        wrapped = BoundaryProjectionSafeguard(env, gate="always", strict=False)
        algo = SHAC(env=wrapped, policy_kwargs={"net_arch": [64, 64]}, policy_optim_kwargs={"lr": 1e-3},
                    vf_kwargs={"net_arch": [64, 64]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                    len_trajectories=8, vf_num_fits=2, vf_fit_num_batches=2)
        for eps in range(2):
            r, pl, vl = algo._learn_episode(eps)
            assert np.isfinite(r) and np.isfinite(pl) and np.isfinite(vl)
        fl = env.last_flags
        guarded = (fl & _lib.FL_GUARD_USED) != 0
        feas = guarded & ((fl & _lib.FL_INFEASIBLE) == 0)
        assert bool(guarded.all()) and int(feas.sum()) > 32
        assert bool((((fl & _lib.FL_NEXT_IN_SET) != 0) | ~feas).all())          # zero safety violations where a safe action exists
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)
