"""GPU: the drop-in classes (envs, safeguards, SHAC) against the reference's golden SHAC episode and the
oracle's safeguarded episode, with the RNG stream replayed draw by draw."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Replay:
    def __init__(self, draws, dev):
        self.draws, self.i, self.dev = list(draws), 0, dev

    def __call__(self, kind, shape):
        k, d = self.draws[self.i]
        self.i += 1
        assert k == kind and tuple(d.shape) == tuple(shape), (k, kind, d.shape, shape)
        return torch.from_numpy(np.asarray(d)).to(self.dev)


@pytest.fixture
def cuda_f64():
    prev_dtype = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda:0")
    yield torch.device("cuda:0")
    torch.set_default_device("cpu")
    torch.set_default_dtype(prev_dtype)
    from safegbpo_b200 import rng
    rng.set_source(None)


def test_shac_episode_vs_reference_golden(cuda_f64, golden_dir):
    """One unsafeguarded SHAC episode (truncation + reset-all inside the horizon): loss, buffer and policy
    gradients equal the reference's (tests/golden/shac_cartpole_unsafe.npz)."""
    from safegbpo_b200 import rng
    from safegbpo_b200.envs.balance_cartpole import BalanceCartPoleEnv
    from safegbpo_b200.learning_algorithms.shac import SHAC
    g = np.load(golden_dir / "shac_cartpole_unsafe.npz")
    keys = sorted(k for k in g.files if k.startswith("draw"))
    draws = [(k.split("_")[1], g[k]) for k in keys]
    env = BalanceCartPoleEnv(num_envs=4, num_steps=3)
    algo = SHAC(env=env, policy_kwargs={"net_arch": [32, 32]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [32, 32, 32]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=6, fused="off")
    algo.policy.load_state_dict({k[len("policy."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("policy.")})
    algo.target_value_function.load_state_dict(
        {k[len("target_vf."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("target_vf.")})
    env.state = torch.from_numpy(g["state0"]).to(cuda_f64)
    obs0 = env.observation
    algo.buffer.reset(obs0, algo.target_value_function(obs0).squeeze(1))
    rng.set_source(Replay(draws, cuda_f64))
    algo.buffer.reset()
    algo.policy_optim.zero_grad()
    avg_reward = algo.collect_trajectories()
    loss = algo.policy_loss()
    loss.backward()
    tol = dict(rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(algo.buffer.observations.tensor.detach().cpu().numpy(), g["observations"], **tol)
    np.testing.assert_allclose(algo.buffer.rewards.tensor.detach().cpu().numpy(), g["rewards"], **tol)
    np.testing.assert_allclose(algo.buffer.values.tensor.detach().cpu().numpy(), g["values"], **tol)
    assert np.array_equal(algo.buffer.terminals.tensor.cpu().numpy(), g["terminals"])
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-11)
    np.testing.assert_allclose(avg_reward, float(g["avg_reward"]), rtol=1e-11)
    for k, p in algo.policy.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), g[f"grad.{k}"], rtol=1e-8, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("sg_name", ["boundary_projection", "ray_mask_orth"])
def test_safeguarded_episode_vs_oracle(cuda_f64, sg_name):
    """SwingUpCartPole + safeguard, H = 8, N = 16: eager drop-in classes vs the oracle restatement."""
    from oracle.dynamics import OracleEnv
    from oracle.safeguard import OracleSafeguard
    from oracle import shac as oshac
    from safegbpo_b200 import rng
    from safegbpo_b200.envs.swingup_cartpole import SwingUpCartPoleEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    from safegbpo_b200.learning_algorithms.shac import SHAC
    N, H = 16, 8
    gen = torch.Generator().manual_seed(5)
    s0 = (torch.rand(N, 4, generator=gen, dtype=torch.float64, device="cpu") * 2 - 1) * torch.tensor(
        [3.0, 3.0, 7.0, 7.0], dtype=torch.float64, device="cpu")
    eps = torch.randn(H, N, 1, generator=gen, dtype=torch.float64, device="cpu")
    us = torch.rand(H, N, 1, 4, generator=gen, dtype=torch.float64, device="cpu")
    # oracle (float64, CPU)
    with torch.device("cpu"):
        oenv = OracleEnv("SwingUpCartPole", N, 1000, rng=lambda kind, shape: next(it_o))
        it_o = iter([u for u in us])
        oenv.state, oenv.scheduled_reset = s0.clone(), False
        kind = "boundary_projection" if sg_name == "boundary_projection" else "ray_mask"
        osg = OracleSafeguard(oenv, kind, zonotopic_approximation=False, gate="always", strict=False)
        pol = oshac.OraclePolicy(5, 1, [32, 32]).double().cpu()
        vf = oshac.OracleValue(5, [32, 32]).double().cpu()

        def env_step(action, t):
            obs, rew, term, trunc = osg.step(action)
            return obs, rew, term | trunc
        obs0 = oenv.observation()
        ref = oshac.rollout_loss(env_step, obs0, pol, vf, H, eps, value0=vf(obs0).squeeze(1))
        ref["loss"].backward()
    # drop-in classes on the GPU
    env = SwingUpCartPoleEnv(num_envs=N, num_steps=1000)
    env.reset()
    env.state = s0.to(cuda_f64)
    cls = BoundaryProjectionSafeguard if sg_name == "boundary_projection" else RayMaskSafeguard
    kw = {} if sg_name == "boundary_projection" else dict(linear_projection=True, zonotopic_approximation=False, passthrough=False)
    wrapped = cls(env, gate="always", strict=False, **kw)
    algo = SHAC(env=wrapped, policy_kwargs={"net_arch": [32, 32]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [32, 32]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=H, fused="off")
    algo.policy.load_state_dict({k: v.to(cuda_f64) for k, v in pol.state_dict().items()})
    algo.target_value_function.load_state_dict({k: v.to(cuda_f64) for k, v in vf.state_dict().items()})
    env.state, env.steps, env._reset_pending = s0.to(cuda_f64), 0, False
    o0 = env.observation
    algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
    draws = []
    for t in range(H):
        draws += [("randn", eps[t].numpy()), ("rand", us[t].numpy())]
    rng.set_source(Replay(draws, cuda_f64))
    algo.buffer.reset()
    algo.collect_trajectories()
    loss = algo.policy_loss()
    loss.backward()
    assert not bool(osg.infeasible.any())
    np.testing.assert_allclose(algo.buffer.observations.tensor.detach().cpu().numpy(), ref["observations"].detach().numpy(),
                               rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=1e-10)
    for (k, p), (k2, p2) in zip(algo.policy.named_parameters(), pol.named_parameters()):
        np.testing.assert_allclose(p.grad.cpu().numpy(), p2.grad.numpy(), rtol=1e-7, atol=1e-10, err_msg=k)


def test_household_zonotopic_ray_mask_through_the_drop_in_classes(cuda_f64):
    """ManageHouseholdEnv + RayMaskSafeguard(zonotopic): program P2 with four free lengths is solved per environment
    inside the step kernel (interior point, csrc/sgbpo_safeguard2.cuh); the wrapper runs and differentiates."""
    from safegbpo_b200 import _lib
    from safegbpo_b200.envs.manage_household import ManageHouseholdEnv
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    N = 48
    torch.manual_seed(2)
    env = ManageHouseholdEnv(num_envs=N, num_steps=720)
    wrapped = RayMaskSafeguard(env, linear_projection=True, zonotopic_approximation=True, passthrough=False,
                               gate="always", strict=False)
    wrapped.reset(seed=4)
    for _ in range(3):
        action = (torch.rand(N, 2) * 2 - 1).requires_grad_(True)
        obs, rew, term, trunc, _ = wrapped.step(action)
        flags = env.last_flags
        feasible = (flags & _lib.FL_INFEASIBLE) == 0
        # (the as-shipped map, Q2, pushes some environments out of the safe set, whose programs are then infeasible)
        assert float(feasible.float().mean()) > 0.5
        a_exec = env.last_exec_action
        # (as shipped the linear radial map lacks the safe_dist factor, Q2: the mapped action may leave the unit box
        #  and the safe set, so only finiteness is asserted here; values and gradients are pinned against the oracle
        #  in test_gpu_step_parity)
        assert bool(torch.isfinite(a_exec[feasible]).all())
        rew[feasible].sum().backward()
        assert bool(torch.isfinite(action.grad[feasible]).all()) and float(action.grad[feasible].abs().sum()) > 0
        env.state = env.state.detach()


@pytest.mark.parametrize("vf_graphs", [False, True], ids=["eager", "graph"])
def test_critic_path_vs_reference_golden(cuda_f64, golden_dir, vf_graphs):
    """TD-lambda targets (csrc/sgbpo_critic.cu), the critic fit (eager torch loop and its CUDA-graph replay with the
    fused clip+Adam launches) and the Polyak update against the REFERENCE's own outputs on the buffer of the golden
    SHAC episode (tests/golden/shac_critic.npz, made by tests/golden/make_golden.py from shac.py:154-251)."""
    from safegbpo_b200.envs.balance_cartpole import BalanceCartPoleEnv
    from safegbpo_b200.learning_algorithms.shac import SHAC
    g, c = np.load(golden_dir / "shac_cartpole_unsafe.npz"), np.load(golden_dir / "shac_critic.npz")
    env = BalanceCartPoleEnv(num_envs=4, num_steps=3)
    algo = SHAC(env=env, policy_kwargs={"net_arch": [32, 32]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [32, 32, 32]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=6, fused="off", td_weight=float(c["td_weight"]), vf_num_fits=int(c["vf_num_fits"]),
                vf_fit_num_batches=int(c["vf_fit_num_batches"]), polyak_target=float(c["polyak_target"]))
    algo.vf_graphs = vf_graphs
    t = lambda a: torch.from_numpy(np.asarray(a)).to(cuda_f64)
    buf = algo.buffer
    buf.observations.tensor, buf.values.tensor, buf.rewards.tensor = t(g["observations"]), t(g["values"]), t(g["rewards"])
    buf.terminals.tensor = t(g["terminals"])
    buf.t = torch.full_like(buf.t, 6)
    algo.value_function.load_state_dict({k[len("vf0."):]: t(c[k]) for k in c.files if k.startswith("vf0.")})
    algo.target_value_function.load_state_dict({k[len("target_vf."):]: t(g[k]) for k in g.files if k.startswith("target_vf.")})
    with torch.no_grad():
        est, obs = algo.calculate_estimated_values()
    np.testing.assert_allclose(est.cpu().numpy(), c["estimated_values"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(obs.cpu().numpy(), c["estimated_observations"], rtol=1e-12, atol=1e-12)
    loss = algo.update_value_function()
    np.testing.assert_allclose(loss, float(c["value_loss"]), rtol=1e-7)
    for k, p in algo.value_function.state_dict().items():
        np.testing.assert_allclose(p.cpu().numpy(), c[f"vf1.{k}"], rtol=1e-7, atol=1e-9, err_msg=k)
    algo.update_target_value_function()
    for k, p in algo.target_value_function.state_dict().items():
        np.testing.assert_allclose(p.cpu().numpy(), c[f"target_vf1.{k}"], rtol=1e-9, atol=1e-11, err_msg=k)
    if vf_graphs:                                   # a second fit replays the captured graph and keeps training
        before = [p.detach().clone() for p in algo.value_function.parameters()]
        loss2 = algo.update_value_function()
        assert algo._vf_graph is not None and algo._vf_graph["graph"] is not None
        assert loss2 < loss and any(not torch.equal(a, b) for a, b in zip(before, algo.value_function.parameters()))
