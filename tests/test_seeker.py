"""NavigateSeeker: per-step safe action set (P5), projection / ray mask on it, collision operator.
CPU: device math compiled for the host vs the oracle.  GPU: the kernels through the C ABI and the drop-in
environment class."""
import numpy as np
import pytest
import torch

import hostmath_util as hm
from seeker_common import collision_case, safeguard_case
from safegbpo_b200 import _lib


@pytest.mark.parametrize("kind,linear", [("boundary_projection", True), ("ray_mask", True), ("ray_mask", False)])
def test_hostmath_safeguard_vs_oracle(kind, linear):
    c = safeguard_case(kind, linear)
    out = hm.step(5, c["consts"], c["s"].numpy(), c["a"].numpy(), c["w"].numpy(), aux=c["aux"].numpy(),
                  cot={k: v.numpy() for k, v in c["cot"].items()})
    ok = c["ref"]["ok"].numpy()
    assert ok.sum() >= 10
    for key in ("a_exec", "s_next", "reward", "gs", "ga"):
        np.testing.assert_allclose(out[key][ok], c["ref"][key].numpy()[ok], rtol=1e-8, atol=1e-9, err_msg=key)
    assert ((out["flags"] & _lib.FL_ACTION_IN_SET) != 0)[ok].all() or kind == "ray_mask"


def test_hostmath_collision_operator_vs_oracle():
    c = collision_case()
    out = hm.step(5, c["consts"], c["s"].numpy(), c["a"].numpy(), c["w"].numpy(), aux=c["aux"].numpy(),
                  cot={k: v.numpy() for k, v in c["cot"].items()})
    assert np.array_equal((out["flags"] & _lib.FL_COLLIDED) != 0, c["ref"]["collided"].numpy())
    assert c["ref"]["collided"].any() and not c["ref"]["collided"].all()
    for key in ("s_next", "reward", "gs", "ga"):
        np.testing.assert_allclose(out[key], c["ref"][key].numpy(), rtol=1e-10, atol=1e-12, err_msg=key)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 2e-4)])
def test_gpu_kernels_vs_oracle(dtype, tol):
    from safegbpo_b200 import ops
    dev = torch.device("cuda:0")
    t = lambda x: x.to(device=dev, dtype=dtype)
    for c in (safeguard_case("boundary_projection", True), safeguard_case("ray_mask", False), collision_case()):
        s, a = t(c["s"]).requires_grad_(True), t(c["a"]).requires_grad_(True)
        s_next, a_exec, obs, rew, flags = ops.safeguarded_step(s, a, t(c["w"]), 5, c["consts"], aux=t(c["aux"]))
        ok = c["ref"].get("ok", torch.ones(s.shape[0], dtype=torch.bool)).to(dev)
        cot = {k: t(v) for k, v in c["cot"].items()}
        func = ((s_next * cot["g_s_next"]).sum(1) + rew * cot["g_reward"])
        if "g_a_exec" in cot:
            func = func + (a_exec * cot["g_a_exec"]).sum(1)
        func[ok].sum().backward()
        okc = ok.cpu().numpy()
        for got, key in ((s_next, "s_next"), (rew, "reward"), (s.grad, "gs"), (a.grad, "ga")):
            np.testing.assert_allclose(got.detach().cpu().double().numpy()[okc], c["ref"][key].numpy()[okc], rtol=tol, atol=tol,
                                       err_msg=key)


@pytest.mark.gpu
def test_gpu_env_class_reset_step_and_action_set():
    """Drop-in NavigateSeekerEnv: rejection-sampled reset invariants, safe_action_set() == oracle P5, and a
    safeguarded step through BoundaryProjectionSafeguard keeps the seeker out of the obstacle."""
    from oracle.dynamics import OracleEnv
    from safegbpo_b200.envs.navigate_seeker import NavigateSeekerEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda:0")
    try:
        torch.manual_seed(0)
        N = 64
        env = NavigateSeekerEnv(num_envs=N, num_steps=400, num_obstacles=1, min_radius=2.0, max_radius=4.0)
        obs, _ = env.reset(seed=3)
        assert obs.shape == (N, 7)
        ball = env.obstacles[0]
        assert not ball.contains(env.state).any() and not ball.contains(env.goal).any()
        assert (torch.linalg.vector_norm(env.goal - env.state, dim=1) >= 8.0).all()
        z = env.safe_action_set()
        with torch.device("cpu"):
            oenv = OracleEnv("NavigateSeeker", N, 400)
            oenv.state, oenv.goal = env.state.cpu(), env.goal.cpu()
            oenv.obstacles = [(ball.center.cpu(), ball.radius.cpu())]
            _, gen = oenv.safe_action_set()
        np.testing.assert_allclose(z.generator.cpu().numpy(), gen.numpy(), rtol=1e-6, atol=1e-8)
        wrapped = BoundaryProjectionSafeguard(env, gate="always", strict=False)
        for _ in range(30):                      # drive straight at the obstacle
            to_obs = ball.center - env.state
            action = to_obs / torch.linalg.vector_norm(to_obs, dim=1, keepdim=True)
            pre = torch.linalg.vector_norm(to_obs, dim=1) - ball.radius
            o, r, term, trunc, _ = wrapped.step(action)
            assert o.shape == (N, 7)
            # P5 bounds the ACTION support by the distance (displacement = DT * a = a / 10) and ignores the
            # noise, so the free step can only enter the obstacle when the noise (|w| <= 0.05 sqrt 2)
            # exceeds the 0.9 * distance the safe action leaves: collided => pre-step distance < 0.08
            assert (pre[env.collided[:, 0]] < 0.08).all()
        dist = torch.linalg.vector_norm(env.state - ball.center, dim=1) - ball.radius
        assert (dist > -0.06).all()              # noise (+-0.05) is not part of the reference's safe set (P5 ignores it)
        assert wrapped.interventions > 0
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)
