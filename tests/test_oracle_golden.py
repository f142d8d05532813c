"""Pin the oracle against the reference's own outputs (tests/golden/*.npz, made by make_golden.py
from the unmodified reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.dynamics import OracleEnv
from oracle.sets import aab_intersects_zonotope
from oracle import shac as oshac

ENVS = ["BalancePendulum", "BalanceCartPole", "SwingUpCartPole", "BalanceQuadrotor", "ManageHousehold"]
TOL = dict(rtol=1e-12, atol=1e-12)


class Replay:
    """Feeds recorded RNG draws to the oracle in order."""

    def __init__(self, draws):
        self.draws, self.i = list(draws), 0

    def __call__(self, kind, shape):
        k, d = self.draws[self.i]
        self.i += 1
        assert k == kind and tuple(d.shape) == tuple(shape), (k, kind, d.shape, shape)
        return torch.from_numpy(np.asarray(d)).clone()


@pytest.mark.parametrize("name", ENVS)
def test_env_matches_reference(name, golden_dir):
    g = np.load(golden_dir / f"env_{name}.npz")
    n_reset = len([k for k in g.files if k.startswith("reset_draw")])
    draws = [("rand", g[f"reset_draw{i}"]) for i in range(n_reset)]
    env = OracleEnv(name, 5, 100, rng=Replay(draws))
    env.reset()
    np.testing.assert_allclose(env.state.numpy(), g["reset_state"], **TOL)
    c, a_mat, b_mat, e_mat = env.linear_dynamics()
    for got, key in [(c, "lin_c"), (a_mat, "lin_A"), (b_mat, "lin_B"), (e_mat, "lin_E")]:
        np.testing.assert_allclose(got.numpy(), g[key], **TOL)
    rc, rg = env.reachable_set()
    np.testing.assert_allclose(rg.numpy(), g["reach_generator"], **TOL)
    gate = aab_intersects_zonotope(env.state_center, env.state_half, rc, rg)
    assert np.array_equal(gate.numpy(), g["projectable"])
    for t in range(3):
        state = torch.from_numpy(g[f"s{t}"]).clone().requires_grad_(True)
        action = torch.from_numpy(g[f"a{t}"]).clone().requires_grad_(True)
        env.state = state
        env.rng = Replay([("rand", g[f"u{t}"])])
        obs, rew, term, trunc = env.step(action)
        np.testing.assert_allclose(env.state.detach().numpy(), g[f"next{t}"], **TOL)
        np.testing.assert_allclose(obs.detach().numpy(), g[f"obs{t}"], **TOL)
        np.testing.assert_allclose(rew.detach().numpy(), g[f"rew{t}"], **TOL)
        func = (obs * torch.from_numpy(g[f"cot_o{t}"])).sum() + (rew * torch.from_numpy(g[f"cot_r{t}"])).sum()
        gs, ga = torch.autograd.grad(func, (state, action))
        np.testing.assert_allclose(gs.numpy(), g[f"gs{t}"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ga.numpy(), g[f"ga{t}"], rtol=1e-10, atol=1e-12)
        env.state = env.state.detach()


def test_shac_episode_matches_reference(golden_dir):
    g = np.load(golden_dir / "shac_cartpole_unsafe.npz")
    n = int(g["n_draws"])
    keys = sorted(k for k in g.files if k.startswith("draw"))
    assert len(keys) == n
    draws = [(k.split("_")[1], g[k]) for k in keys]
    # split the recorded stream: policy draws (randn) vs env draws (rand), order preserved per consumer
    eps = torch.from_numpy(np.stack([d for k, d in draws if k == "randn"]))
    env = OracleEnv("BalanceCartPole", 4, 3, rng=Replay([(k, d) for k, d in draws if k == "rand"]))
    env.state = torch.from_numpy(g["state0"]).clone()
    env.scheduled_reset = False
    policy = oshac.OraclePolicy(5, 1, [32, 32]).double()
    vf = oshac.OracleValue(5, [32, 32, 32]).double()
    policy.load_state_dict({k[len("policy."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("policy.")})
    vf.load_state_dict({k[len("target_vf."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("target_vf.")})

    def env_step(action, t):
        obs, rew, term, trunc = env.step(action)
        return obs, rew, term | trunc

    obs0 = env.observation()
    out = oshac.rollout_loss(env_step, obs0, policy, vf, 6, eps, value0=vf(obs0).squeeze(1))
    np.testing.assert_allclose(out["observations"].detach().numpy(), g["observations"], **TOL)
    np.testing.assert_allclose(out["rewards"].detach().numpy(), g["rewards"], **TOL)
    np.testing.assert_allclose(out["values"].detach().numpy(), g["values"], **TOL)
    assert np.array_equal(out["terminals"].numpy(), g["terminals"])
    np.testing.assert_allclose(out["loss"].detach().numpy(), g["loss"], rtol=1e-12)
    out["loss"].backward()
    for k, p in policy.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g[f"grad.{k}"], rtol=1e-9, atol=1e-13)
