"""GPU: the consumers of the forward kernels outside SHAC (learning_algorithms/consumers.py): policy evaluation
(reference logger.py:91-131) and the model-free rollout collection (ppo.py:99-120), whole-horizon engine vs per-step loop."""
import numpy as np
import pytest
import torch

from test_gpu_fused_rollout import Recorder, Replay, cuda_default  # noqa: F401

pytestmark = pytest.mark.gpu


def _make(dtype, fused, dev):
    from safegbpo_b200.envs.balance_cartpole import BalanceCartPoleEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.learning_algorithms.components.policy import Policy
    from safegbpo_b200.learning_algorithms.components.value_function import ValueFunction
    from safegbpo_b200.learning_algorithms.consumers import RolloutCollector
    torch.set_default_dtype(dtype)
    torch.manual_seed(4)
    env = BoundaryProjectionSafeguard(BalanceCartPoleEnv(num_envs=200, num_steps=7), gate="intended", strict=False)
    policy, vf = Policy(env.obs_dim, env.action_dim, net_arch=[128, 128]), ValueFunction(env.obs_dim, net_arch=[128, 128, 128])
    return env, policy, vf, RolloutCollector(env, policy, vf, len_trajectories=12, fused=fused)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_collector_engine_equals_per_step_loop(cuda_default, dtype):
    from safegbpo_b200 import rng
    dev = cuda_default
    env_a, pol_a, vf_a, col_a = _make(dtype, "off", dev)
    env_b, pol_b, vf_b, col_b = _make(dtype, "on", dev)
    pol_b.load_state_dict(pol_a.state_dict())
    vf_b.load_state_dict(vf_a.state_dict())
    half = env_a.state_set.generator[0].diagonal()
    s0 = env_a.state_set.center[0] + (torch.rand(200, 4) * 2 - 1) * half * 0.6
    for env, col, vf in ((env_a, col_a, vf_a), (env_b, col_b, vf_b)):
        env.env.state, env.env.steps, env.env._reset_pending = s0.clone(), 2, False
        o0 = env.env.observation
        with torch.no_grad():
            col.buffer.reset(o0, vf(o0).squeeze(1))
    rec = Recorder()
    rng.set_source(rec)
    r_a = col_a.collect()
    rng.set_source(Replay(rec.draws, dev))
    r_b = col_b.collect()
    rng.set_source(None)
    assert col_b._fused_engine is not None and col_b._fused_engine.forward_only
    tol = dict(rtol=1e-9, atol=1e-10) if dtype == torch.float64 else dict(rtol=2e-4, atol=2e-4)
    H = 12
    for name in ("observations", "actions", "rewards", "values", "log_probs"):
        a, b = getattr(col_a.buffer, name).tensor, getattr(col_b.buffer, name).tensor
        rows = H if name in ("actions", "rewards", "log_probs") else H + 1
        np.testing.assert_allclose(b[:rows].cpu().numpy(), a[:rows].cpu().numpy(), err_msg=name, **tol)
    assert torch.equal(col_a.buffer.terminals.tensor[:H + 1], col_b.buffer.terminals.tensor[:H + 1])
    np.testing.assert_allclose(r_b, r_a, rtol=tol["rtol"])
    assert env_a.interventions == env_b.interventions


def test_evaluate_policy_matches_manual_loop(cuda_default):
    from safegbpo_b200.learning_algorithms.consumers import evaluate_policy
    torch.set_default_dtype(torch.float64)
    env, policy, _, _ = _make(torch.float64, "off", cuda_default)
    got = evaluate_policy(env, policy)
    obs, _ = env.eval_reset()
    total, steps, terminal = 0.0, 0, False
    with torch.no_grad():
        while not terminal:                       # logger.py:104-117
            obs, reward, terminated, truncated, _ = env.step(policy.predict(obs, deterministic=True))
            terminal = bool((terminated | truncated)[0])
            total += float(reward.sum())
            steps += 1
    assert steps == 7
    np.testing.assert_allclose(got, total / env.num_envs / steps, rtol=1e-12)


def test_log_prob_matches_the_change_of_variables():
    """Policy.log_prob (reference policy.py:112-141): inside (-1, 1) the tanh change of variables, at +-1 the Gaussian tail."""
    from safegbpo_b200.learning_algorithms.components.policy import Policy
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        with torch.device("cpu"), torch.no_grad():
            torch.manual_seed(0)
            pol = Policy(3, 2, net_arch=[16, 16])
            x = torch.randn(64, 3)
            mean, std = pol.model(x), torch.exp(pol.log_std)
            u = mean + std * torch.randn(64, 2)
            a = torch.tanh(u)
            ref = (torch.distributions.Normal(mean, std).log_prob(u) - torch.log(1 - a ** 2)).sum(1)
            np.testing.assert_allclose(pol.log_prob(a, x).numpy(), ref.numpy(), rtol=1e-8, atol=1e-8)
            sat = torch.ones(64, 2)
            tiny = torch.finfo(torch.float64).eps
            tail = torch.log(1 - torch.distributions.Normal(mean, std).cdf(torch.atanh(sat.clamp(max=1 - tiny))) + tiny).sum(1)
            np.testing.assert_allclose(pol.log_prob(sat, x).numpy(), tail.numpy(), rtol=1e-8, atol=1e-8)
    finally:
        torch.set_default_dtype(prev)
