"""GPU: the CUDA kernels, called through the C ABI, against the oracle and the reference goldens.

Tolerances: float64 kernels must agree with the float64 oracle to 1e-9 relative (same algorithm, different
evaluation order); float32 kernels to 1e-5 relative (the tolerance BASELINE.json's north_star states),
checked on the environments whose float32 run takes the same branches as the float64 oracle."""
import numpy as np
import pytest
import torch

from common import ENV_IDS, make_case, oracle_step, p2_center_unique, shared_for
from safegbpo_b200 import _lib, ops

pytestmark = pytest.mark.gpu

CASES = [
    ("BalanceCartPole", "none", {}), ("BalancePendulum", "none", {}), ("BalanceQuadrotor", "none", {}),
    ("ManageHousehold", "none", {}), ("SwingUpCartPole", "none", {}),
    ("BalanceCartPole", "boundary_projection", {}), ("SwingUpCartPole", "boundary_projection", {}),
    ("BalanceCartPole", "boundary_projection", {"gate": "reference"}),
    ("SwingUpCartPole", "ray_mask", {"zonotopic": False}), ("SwingUpCartPole", "ray_mask", {"zonotopic": False, "linear": False}),
    ("SwingUpCartPole", "ray_mask", {"zonotopic": True}), ("BalanceCartPole", "ray_mask", {"zonotopic": True, "linear": False}),
    ("SwingUpCartPole", "ray_mask", {"zonotopic": False, "passthrough": True}),
    ("ManageHousehold", "boundary_projection", {}), ("ManageHousehold", "ray_mask", {"zonotopic": False}),
    ("ManageHousehold", "ray_mask", {"zonotopic": True}),
    ("BalancePendulum", "boundary_projection", {}), ("BalancePendulum", "ray_mask", {"zonotopic": False}),
    ("BalancePendulum", "ray_mask", {"zonotopic": True}), ("BalancePendulum", "ray_mask", {"zonotopic": True, "gate": "reference"}),
]
IDS = [f"{c[0]}-{c[1]}-{'-'.join(f'{k}={v}' for k, v in c[2].items())}" for c in CASES]


def run_kernel(case, env_name, dtype, dev):
    t = lambda x: None if x is None else x.to(device=dev, dtype=dtype)
    s = t(case["s"]).requires_grad_(True)
    a = t(case["a"]).requires_grad_(True)
    out = ops.safeguarded_step(s, a, t(case["w"]), ENV_IDS[env_name], case["consts"], aux=t(case["aux"]),
                               dirs=t(case["dirs"]), shared=shared_for(case["env"], 1))
    return s, a, out


@pytest.mark.parametrize("env_name,kind,opts", CASES, ids=IDS)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_step_fwd_bwd_vs_oracle(env_name, kind, opts, dtype):
    dev = torch.device("cuda:0")
    zono = kind == "ray_mask" and opts.get("zonotopic", True)
    N = 12 if zono else 64
    case = make_case(env_name, kind, N, seed=11, **opts)
    ref = oracle_step(case)
    s, a, (s_next, a_exec, obs, reward, flags) = run_kernel(case, env_name, dtype, dev)
    ok = ref["ok"].to(dev)
    if env_name == "ManageHousehold" and zono:
        # P2's optimal centre is not unique for some direction draws: only the unique ones are comparable
        ok = ok & torch.from_numpy(p2_center_unique(case)).to(dev)
        assert int(ok.sum()) >= 3
    flags_c = flags.cpu().numpy()
    if dtype == torch.float32:
        # fp32 may take a different branch (clamp / active row) when within rounding of a switch: compare
        # the environments whose flag word equals the float64 kernel's
        _, _, out64 = run_kernel(case, env_name, torch.float64, dev)
        same = (out64[4] == flags)
        assert same.float().mean() > 0.9
        ok = ok & same
    cot = {k: v.to(device=dev, dtype=dtype) for k, v in case["cot"].items()}
    func = ((s_next * cot["g_s_next"]).sum(1) + (a_exec * cot["g_a_exec"]).sum(1) + (obs * cot["g_obs"]).sum(1)
            + reward * cot["g_reward"])[ok].sum()
    func.backward()
    torch.cuda.synchronize()
    okc = ok.cpu().numpy()
    if dtype == torch.float64:
        # zonotopic ray mask: the ORACLE's P2 is an SLSQP + Newton-polish solve, accurate to ~1e-6
        tol = dict(rtol=1e-5, atol=2e-6) if zono else dict(rtol=1e-9, atol=1e-10)
        gtol = dict(rtol=1e-4, atol=1e-5) if zono else dict(rtol=1e-8, atol=1e-9)
    else:
        tol = dict(rtol=1e-5, atol=2e-5)          # float32: 1e-5 relative (north_star), absolute floor for O(10) states
        gtol = dict(rtol=2e-4, atol=2e-4)         # first derivatives of fp32 kinks: looser, stated here
        if kind == "ray_mask":
            # the ray-mask maps divide by the safe chord length hi - lo (ray_mask.py:109-111, Q2): fp32 cancellation
            # in that difference is amplified by 1 / chord, so the float32 tolerance is 2e-4 / 2e-3 (grads) here
            tol, gtol = dict(rtol=2e-4, atol=2e-4), dict(rtol=2e-3, atol=2e-3)
    for name, got in (("a_exec", a_exec), ("s_next", s_next), ("obs", obs), ("reward", reward)):
        np.testing.assert_allclose(got.detach().cpu().double().numpy()[okc], ref[name].numpy()[okc], err_msg=name, **tol)
    np.testing.assert_allclose(s.grad.cpu().double().numpy()[okc], ref["gs"].numpy()[okc], err_msg="gs", **gtol)
    np.testing.assert_allclose(a.grad.cpu().double().numpy()[okc], ref["ga"].numpy()[okc], err_msg="ga", **gtol)
    if kind != "none" and dtype == torch.float64:
        assert np.array_equal((flags_c & _lib.FL_PROJECTABLE) != 0, ref["projectable"].numpy())
        guard = (flags_c & _lib.FL_GUARD_USED) != 0
        assert np.array_equal(((flags_c & _lib.FL_INFEASIBLE) != 0)[guard], ref["infeasible"].numpy()[guard])
        # verdicts: every feasible guarded action is inside its sets (zero safety violations)
        good = guard & ((flags_c & _lib.FL_INFEASIBLE) == 0)
        env = case["env"]
        if env.state_constrained and kind == "boundary_projection":
            assert ((flags_c & _lib.FL_NEXT_IN_SET) != 0)[good].all()
        if env.action_constrained and kind == "boundary_projection":
            assert ((flags_c & _lib.FL_ACTION_IN_SET) != 0)[good].all()


@pytest.mark.parametrize("env_name", ["BalancePendulum", "BalanceCartPole", "SwingUpCartPole", "BalanceQuadrotor", "ManageHousehold"])
def test_against_reference_golden(env_name, golden_dir):
    """Kernel vs the REFERENCE's own outputs (tests/golden/env_*.npz, made from the unmodified reference)."""
    g = np.load(golden_dir / f"env_{env_name}.npz")
    dev = torch.device("cuda:0")
    from safegbpo_b200.consts import build_consts, BOXES, ENV_SIM
    k = build_consts(env_name, _lib.SG_NONE)
    eid = ENV_IDS[env_name]
    s = torch.from_numpy(g["reset_state"]).to(dev)
    c, A, B, E = ops.linearise(s, eid, k)
    for got, key in ((c, "lin_c"), (A, "lin_A"), (B, "lin_B"), (E, "lin_E")):
        np.testing.assert_allclose(got.cpu().numpy(), g[key], rtol=1e-12, atol=1e-12, err_msg=key)
    bc, bh, nc, nh = (np.asarray(x, dtype=np.float64) for x in BOXES[ENV_SIM[env_name]])
    from oracle.dynamics import OracleEnv
    oenv = OracleEnv(env_name, 5, 100)
    for t in range(3):
        st = torch.from_numpy(g[f"s{t}"]).to(dev).requires_grad_(True)
        ac = torch.from_numpy(g[f"a{t}"]).to(dev).requires_grad_(True)
        noise = nc + nh * (2 * g[f"u{t}"][:, 0, :] - 1)
        step_idx = int(g[f"steps{t}"])
        if env_name == "ManageHousehold":
            noise[:, 1] = float(oenv.t_out[step_idx - 1])
        aux = torch.from_numpy(g[f"goal{t}"]).to(dev) if f"goal{t}" in g.files and env_name == "BalanceQuadrotor" else None
        s_next, a_exec, obs, rew, flags = ops.safeguarded_step(st, ac, torch.from_numpy(noise).to(dev), eid, k, aux=aux,
                                                               shared=shared_for(oenv, step_idx))
        np.testing.assert_allclose(s_next.detach().cpu().numpy(), g[f"next{t}"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(obs.detach().cpu().numpy(), g[f"obs{t}"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rew.detach().cpu().numpy(), g[f"rew{t}"], rtol=1e-12, atol=1e-12)
        func = (obs * torch.from_numpy(g[f"cot_o{t}"]).to(dev)).sum() + (rew * torch.from_numpy(g[f"cot_r{t}"]).to(dev)).sum()
        func.backward()
        np.testing.assert_allclose(st.grad.cpu().numpy(), g[f"gs{t}"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(ac.grad.cpu().numpy(), g[f"ga{t}"], rtol=1e-9, atol=1e-11)


def test_empty_batch_and_bad_args():
    dev = torch.device("cuda:0")
    from safegbpo_b200.consts import build_consts
    k = build_consts("BalanceCartPole", _lib.SG_NONE)
    z = lambda *s: torch.zeros(*s, device=dev)
    out = ops.safeguarded_step(z(0, 4), z(0, 1), z(0, 4), 1, k)
    assert out[0].shape == (0, 4) and out[4].shape == (0,)
    with pytest.raises(_lib.SgbpoError):
        ops.safeguarded_step(z(3, 4), z(3, 2), z(3, 4), 1, k)
    with pytest.raises(_lib.SgbpoError):
        ops.safeguarded_step(z(3, 4).cpu(), z(3, 1).cpu(), z(3, 4).cpu(), 1, k)


def test_large_batch_properties():
    """BASELINE full size (N = 4096 .. 65536): size-independent properties of the projection."""
    dev = torch.device("cuda:0")
    for N in (4096, 65536):
        case = make_case("BalanceCartPole", "boundary_projection", 8, seed=3)
        g = torch.Generator().manual_seed(N)
        s = ((torch.rand(N, 4, generator=g, dtype=torch.float64) * 2 - 1) * torch.tensor([10.0, 3.1, 10.0, 10.0], dtype=torch.float64)).to(dev)
        a = (torch.rand(N, 1, generator=g, dtype=torch.float64) * 2 - 1).to(dev)
        w = torch.zeros(N, 4, dtype=torch.float64, device=dev)
        s_next, a1, obs, rew, fl = ops.safeguarded_step(s, a, w, 1, case["consts"])
        _, a2, _, _, fl2 = ops.safeguarded_step(s, a1.detach(), w, 1, case["consts"])
        feas = (fl & _lib.FL_INFEASIBLE) == 0
        assert torch.allclose(a1[feas], a2[feas], rtol=0, atol=1e-12)            # projection is idempotent
        assert (((fl & _lib.FL_NEXT_IN_SET) != 0) | ~feas).all()                 # zero violations where feasible
        assert (a1[feas].abs() <= 1 + 1e-12).all()
        # (an empty safe interval has no projection: NaN action and next state, flagged FL_INFEASIBLE, like the m = 2 path)
        assert ((s_next[feas] - torch.tensor([10.0, 3.141592653589793, 10.0, 10.0], device=dev, dtype=torch.float64)) <= 1e-12).all()
        assert bool(a1[~feas].isnan().all())
        inside = feas & (((fl & _lib.FL_INTERVENED) == 0))
        assert torch.equal(a1[inside], a[inside])                                # already-safe actions pass unchanged


def _quadrotor_p1_golden(golden_dir):
    g = np.load(golden_dir / "quadrotor_p1_oracle.npz")
    case = make_case("BalanceQuadrotor", "boundary_projection", 2, seed=7, gate="always")      # constants only
    assert case["consts"].state_model == 3 and case["consts"].n_gen == 18 and case["consts"].n_null == 12
    return g, case["consts"]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_quadrotor_lifted_projection_vs_oracle_golden(golden_dir, dtype):
    """P1 on BalanceQuadrotor's 6 x 24 safe state set (per-environment interior point, csrc/sgbpo_lifted.cuh) against the
    oracle's lifted QP on states near the RCI boundary and actions outside the safe action set
    (tests/golden/quadrotor_p1_oracle.npz, made by make_quadrotor_p1.py): projected action, next state, observation,
    reward and the gradients w.r.t. state and action."""
    dev = torch.device("cuda:0")
    g, consts = _quadrotor_p1_golden(golden_dir)
    t = lambda x: torch.from_numpy(x).to(device=dev, dtype=dtype)
    s, a = t(g["s"]).requires_grad_(True), t(g["a"]).requires_grad_(True)
    s_next, a_exec, obs, rew, fl = ops.safeguarded_step(s, a, t(g["w"]), ENV_IDS["BalanceQuadrotor"], consts, aux=t(g["aux"]))
    ok = torch.from_numpy(g["ok"]).to(dev)
    assert int(ok.sum()) >= 24
    assert not bool(((fl & _lib.FL_INFEASIBLE) != 0)[ok].any())       # (the oracle's phase 1 misses a few feasible programs)
    ((s_next * t(g["cot_g_s_next"])).sum(1) + (a_exec * t(g["cot_g_a_exec"])).sum(1) + (obs * t(g["cot_g_obs"])).sum(1)
     + rew * t(g["cot_g_reward"]))[ok].sum().backward()
    okc = g["ok"]
    moved = np.abs(g["ref_a_exec"] - g["a"]).max(1) > 1e-6
    assert int((moved & okc).sum()) >= 4                                 # the projection really binds in the fixture
    tol, gtol = (dict(rtol=1e-6, atol=1e-7), dict(rtol=2e-5, atol=2e-5)) if dtype == torch.float64 else \
                (dict(rtol=2e-5, atol=2e-5), dict(rtol=2e-3, atol=2e-3))
    for got, key in ((a_exec, "a_exec"), (s_next, "s_next"), (obs, "obs"), (rew, "reward")):
        np.testing.assert_allclose(got.detach().cpu().double().numpy()[okc], g["ref_" + key][okc], err_msg=key, **tol)
    np.testing.assert_allclose(s.grad.cpu().double().numpy()[okc], g["ref_gs"][okc], err_msg="gs", **gtol)
    np.testing.assert_allclose(a.grad.cpu().double().numpy()[okc], g["ref_ga"][okc], err_msg="ga", **gtol)
    # already-safe actions pass through exactly
    same = ~moved & okc
    assert np.array_equal(a_exec.detach().cpu().double().numpy()[same], t(g["a"]).cpu().double().numpy()[same])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_quadrotor_orthogonal_ray_mask_vs_oracle_golden(golden_dir, dtype):
    """Orthogonal ray mask (P1 then P4, ray_mask.py:268-341) on BalanceQuadrotor's lifted program against the oracle
    fixture tests/golden/quadrotor_rm_orth_oracle.npz (make_quadrotor_p1.py ray_mask)."""
    dev = torch.device("cuda:0")
    g = np.load(golden_dir / "quadrotor_rm_orth_oracle.npz")
    consts = make_case("BalanceQuadrotor", "ray_mask", 2, seed=7, gate="always", zonotopic=False)["consts"]
    assert consts.state_model == 3
    t = lambda x: torch.from_numpy(x).to(device=dev, dtype=dtype)
    s, a = t(g["s"]).requires_grad_(True), t(g["a"]).requires_grad_(True)
    s_next, a_exec, obs, rew, fl = ops.safeguarded_step(s, a, t(g["w"]), ENV_IDS["BalanceQuadrotor"], consts, aux=t(g["aux"]))
    okc = g["ok"]
    ok = torch.from_numpy(okc).to(dev)
    moved = np.abs(g["ref_a_exec"] - g["a"]).max(1) > 1e-6
    assert int(okc.sum()) >= 16 and int((moved & okc).sum()) >= 5
    assert not bool(((fl & _lib.FL_INFEASIBLE) != 0)[ok].any())
    ((s_next * t(g["cot_g_s_next"])).sum(1) + (a_exec * t(g["cot_g_a_exec"])).sum(1) + (obs * t(g["cot_g_obs"])).sum(1)
     + rew * t(g["cot_g_reward"]))[ok].sum().backward()
    tol, gtol = (dict(rtol=5e-6, atol=2e-7), dict(rtol=1e-4, atol=1e-4)) if dtype == torch.float64 else \
                (dict(rtol=5e-5, atol=5e-5), dict(rtol=5e-3, atol=5e-3))
    for got, key in ((a_exec, "a_exec"), (s_next, "s_next"), (obs, "obs"), (rew, "reward")):
        np.testing.assert_allclose(got.detach().cpu().double().numpy()[okc], g["ref_" + key][okc], err_msg=key, **tol)
    np.testing.assert_allclose(s.grad.cpu().double().numpy()[okc], g["ref_gs"][okc], err_msg="gs", **gtol)
    np.testing.assert_allclose(a.grad.cpu().double().numpy()[okc], g["ref_ga"][okc], err_msg="ga", **gtol)


def test_quadrotor_gate_and_unsupported_programs():
    """Under the shipped gate (Q1) RCI states are projectable and the raw action is executed (no solve); far outside the
    feasible box the program is infeasible and says so; the ray-mask programs of this set are not built and say so."""
    dev = torch.device("cuda:0")
    N = 16
    case = make_case("BalanceQuadrotor", "boundary_projection", N, seed=5, gate="reference")
    env = case["env"]
    env.reset(seed=5)
    s0, aux = env.state.detach().clone().to(dev), env.goal.clone().to(dev)
    a, w = case["a"].to(dev), case["w"].to(dev)
    out = ops.safeguarded_step(s0, a, w, ENV_IDS["BalanceQuadrotor"], case["consts"], aux=aux)
    fl = out[4]
    assert bool(((fl & _lib.FL_PROJECTABLE) != 0).all()) and not bool(((fl & _lib.FL_GUARD_USED) != 0).any())
    assert torch.equal(out[1], a)
    far = s0.clone()
    far[:, 0] = 50.0
    out = ops.safeguarded_step(far, a, w, ENV_IDS["BalanceQuadrotor"], case["consts"], aux=aux)
    assert bool(((out[4] & _lib.FL_INFEASIBLE) != 0).all()) and bool(out[1].isnan().all())
    from safegbpo_b200.envs.balance_quadrotor import BalanceQuadrotorEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device(dev)
    try:
        qenv = BalanceQuadrotorEnv(num_envs=N, num_steps=240)
        wrapped = BoundaryProjectionSafeguard(qenv)
        wrapped.reset(seed=1)
        wrapped.step(torch.zeros(N, 2))
        qenv.state = far
        with pytest.raises(ValueError):
            wrapped.step(torch.zeros(N, 2))
        qenv2 = BalanceQuadrotorEnv(num_envs=N, num_steps=240)
        rm = RayMaskSafeguard(qenv2, linear_projection=True, zonotopic_approximation=True, passthrough=False)
        rm.reset(seed=1)
        assert rm.consts().state_model == 2
        qenv2.state = far
        with pytest.raises(NotImplementedError):
            rm.step(torch.zeros(N, 2))
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)


@pytest.mark.parametrize("env_name,kind,N", [
    ("SwingUpCartPole", "boundary_projection", 16384), ("SwingUpCartPole", "ray_mask", 16384),
    ("BalanceQuadrotor", "boundary_projection", 32768), ("ManageHousehold", "boundary_projection", 65536),
    ("BalancePendulum", "ray_mask", 65536)])
def test_baseline_config_sizes_properties(env_name, kind, N):
    """The other BASELINE.json configurations at their full batch sizes (float64, per-step kernels through the C ABI):
    size-independent properties -- the result does not depend on the batch an environment sits in (the first 64
    environments of the big launch equal a 64-environment launch bit for bit), the projection is idempotent, no
    feasible guarded action leaves its sets, already-safe actions pass unchanged, gradients are finite."""
    dev = torch.device("cuda:0")
    opts = {"zonotopic": False} if kind == "ray_mask" and env_name != "BalancePendulum" else {}
    gate = "reference" if env_name == "BalanceQuadrotor" else "always"
    case = make_case(env_name, kind, 8, seed=3, gate=gate, **opts)
    env = case["env"]
    g = torch.Generator().manual_seed(N)
    c = torch.tensor(env.sim.state_center, dtype=torch.float64)
    h = torch.tensor(env.sim.state_half, dtype=torch.float64)
    spread = 0.08 if env_name in ("BalancePendulum", "BalanceQuadrotor") else 0.95      # RCI sets are a small part of the box
    s = (c + (torch.rand(N, env.state_dim, generator=g, dtype=torch.float64) * 2 - 1) * h * spread).to(dev)
    a = (torch.rand(N, env.action_dim, generator=g, dtype=torch.float64) * 2 - 1).to(dev)
    w = torch.zeros(N, env.state_dim, dtype=torch.float64, device=dev)
    aux = None if case["aux"] is None else s[:, :2].clone()
    dirs = None
    if case["dirs"] is not None:
        d = torch.rand(N, env.action_dim, 2 * env.action_dim, generator=g, dtype=torch.float64) * 2 - 1
        dirs = (d / torch.linalg.vector_norm(d, dim=1, keepdim=True)).to(dev)
    eid, k, sh = ENV_IDS[env_name], case["consts"], shared_for(env, 1)
    sr, ar = s.clone().requires_grad_(True), a.clone().requires_grad_(True)
    s_next, a1, obs, rew, fl = ops.safeguarded_step(sr, ar, w, eid, k, aux=aux, dirs=dirs, shared=sh)
    small = ops.safeguarded_step(s[:64], a[:64], w[:64], eid, k, aux=None if aux is None else aux[:64],
                                 dirs=None if dirs is None else dirs[:64], shared=sh)
    for big, sm in zip((s_next, a1, obs, rew, fl), small):
        if big.is_floating_point():       # (infeasible programs give NaN rows in both launches)
            assert torch.equal(torch.nan_to_num(big[:64].detach(), nan=12345.0), torch.nan_to_num(sm, nan=12345.0)), "result depends on the batch"
        else:
            assert torch.equal(big[:64], sm), "result depends on the batch"
    feas = ((fl & _lib.FL_INFEASIBLE) == 0) & ((fl & _lib.FL_UNREDUCED) == 0)
    guard = (fl & _lib.FL_GUARD_USED) != 0
    assert float(feas.float().mean()) > 0.5
    if kind == "boundary_projection":
        _, a2, _, _, _ = ops.safeguarded_step(s, a1.detach(), w, eid, k, aux=aux, dirs=dirs, shared=sh)
        assert torch.allclose(a1[feas].detach(), a2[feas], rtol=0, atol=1e-10)                       # idempotent
        good = feas & guard
        if env.state_constrained and k.state_model not in (2, 3):
            assert bool((((fl & _lib.FL_NEXT_IN_SET) != 0) | ~good).all())                         # zero violations
        if env.action_constrained:
            assert bool((((fl & _lib.FL_ACTION_IN_SET) != 0) | ~good).all())
        untouched = feas & ((fl & _lib.FL_INTERVENED) == 0) & ~guard
        assert torch.equal(a1[untouched].detach(), a[untouched])
    (rew[feas].sum() + s_next[feas].sum()).backward()
    assert bool(torch.isfinite(sr.grad[feas]).all()) and bool(torch.isfinite(ar.grad[feas]).all())
