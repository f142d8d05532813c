"""Generate golden vectors by running the UNMODIFIED reference (/root/reference/src) in the build
container under the import stubs of oracle/stubs.  The reference cannot travel to the GPU box, so its
outputs are committed as small .npz fixtures next to this script.

    python tests/golden/make_golden.py

Recorded (float64, CPU), for BalancePendulum / BalanceCartPole / SwingUpCartPole / BalanceQuadrotor /
ManageHousehold:
  * reset state for a seed, ``linear_dynamics()`` (c, A, B, E), ``reachable_set()`` and the gate verdict
    ``state_set.intersects(reachable_set)``                       (simulator.py:206-243, box.py:139-154)
  * three ``step`` calls with fixed actions: the noise drawn, next state, observation, reward and the
    gradients of a fixed linear functional of (obs, reward) w.r.t. (state, action)  (simulator.py:82-108)
and one unsafeguarded SHAC episode (shac.py:96-147) on BalanceCartPole with a truncation inside the
horizon (exercises Q6 reset-all and the terminal-aware discounting), with every RNG draw captured.
The CvxpyLayer solves are NOT recordable here (cvxpy stack absent): see oracle/__init__.py.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT / "oracle" / "stubs"), "/root/reference", "/root/reference/src"]

torch.set_default_dtype(torch.float64)

DRAWS = []
_orig_rand = torch.rand


def _rec_rand(*a, **k):
    out = _orig_rand(*a, **k)
    DRAWS.append(("rand", out.detach().clone()))
    return out


torch.rand = _rec_rand
import torch.distributions.normal as _tdn  # noqa: E402

_orig_sn = _tdn._standard_normal


def _rec_sn(shape, dtype, device):
    out = _orig_sn(shape, dtype, device)
    DRAWS.append(("randn", out.detach().clone()))
    return out


_tdn._standard_normal = _rec_sn


def env_golden(cls, name, seed, n_envs=5, n_steps=100):
    env = cls(num_envs=n_envs, num_steps=n_steps)
    DRAWS.clear()
    env.reset(seed=seed)
    out = {"reset_state": env.state.numpy().copy()}
    for i, (_, d) in enumerate(DRAWS):
        out[f"reset_draw{i}"] = d.numpy()
    c, a_mat, b_mat, e_mat = env.linear_dynamics()
    out.update(lin_c=c.numpy(), lin_A=a_mat.numpy(), lin_B=b_mat.numpy(), lin_E=e_mat.numpy())
    reach = env.reachable_set()
    out["reach_center"] = reach.center.numpy()
    out["reach_generator"] = reach.generator.numpy()
    out["projectable"] = env.state_set.intersects(reach).numpy()
    gen = torch.Generator().manual_seed(1000 + seed)
    for t in range(3):
        state = env.state.detach().clone().requires_grad_(True)
        env.state = state
        action = (_orig_rand(n_envs, env.action_dim, generator=gen) * 2 - 1).requires_grad_(True)
        cot_o = _orig_rand(n_envs, env.obs_dim, generator=gen)
        cot_r = _orig_rand(n_envs, generator=gen)
        DRAWS.clear()
        obs, rew, term, trunc, _ = env.step(action)
        noise = DRAWS[0][1]
        functional = (obs * cot_o).sum() + (rew * cot_r).sum()
        g_s, g_a = torch.autograd.grad(functional, (state, action))
        out.update({f"s{t}": state.detach().numpy(), f"a{t}": action.detach().numpy(), f"u{t}": noise.numpy(),
                    f"cot_o{t}": cot_o.numpy(), f"cot_r{t}": cot_r.numpy(),
                    f"next{t}": env.state.detach().numpy(), f"obs{t}": obs.detach().numpy(),
                    f"rew{t}": rew.detach().numpy(), f"gs{t}": g_s.numpy(), f"ga{t}": g_a.numpy(),
                    f"steps{t}": np.array(env.steps)})
        if hasattr(env, "goal"):
            out[f"goal{t}"] = env.goal.detach().numpy()
        env.state = env.state.detach()
    np.savez_compressed(HERE / f"env_{name}.npz", **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if k.startswith(("lin", "proj"))})


def shac_golden():
    from envs.balance_cartpole import BalanceCartPoleEnv
    from learning_algorithms.shac import SHAC
    torch.manual_seed(7)
    env = BalanceCartPoleEnv(num_envs=4, num_steps=3)
    DRAWS.clear()
    algo = SHAC(env=env, policy_kwargs={"net_arch": [32, 32]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [32, 32, 32]}, vf_optim_kwargs={"lr": 1e-3},
                regularisation_coefficient=0.1, len_trajectories=6)
    out = {}
    for k, v in algo.policy.state_dict().items():
        out[f"policy.{k}"] = v.numpy().copy()
    for k, v in algo.target_value_function.state_dict().items():
        out[f"target_vf.{k}"] = v.numpy().copy()
    out["state0"] = env.state.numpy().copy()
    n_init = len(DRAWS)
    # one episode exactly as SHAC._learn_episode does up to the backward pass (shac.py:87-90, 131-147)
    algo.buffer.reset()
    algo.policy_optim.zero_grad()
    avg_reward = algo.collect_trajectories()
    buf = algo.buffer
    exponent = torch.arange(algo.len_trajectories + 2).view(-1, 1).repeat(1, env.num_envs)
    for end, e in buf.terminals.nonzero():
        exponent[end + 1:, e] -= end + 1
    discount = algo.GAMMA ** exponent
    values = torch.where(buf.terminals.tensor, buf.values.tensor, buf.rewards.tensor)
    norm = env.num_envs * algo.len_trajectories + buf.terminals.count_nonzero()
    loss = -(discount * values).sum() / norm
    loss.backward()
    out["avg_reward"] = np.array(avg_reward)
    out["loss"] = loss.detach().numpy()
    for k, p in algo.policy.named_parameters():
        out[f"grad.{k}"] = p.grad.numpy().copy()
    out["observations"] = buf.observations.tensor.detach().numpy()
    out["values"] = buf.values.tensor.detach().numpy()
    out["rewards"] = buf.rewards.tensor.detach().numpy()
    out["actions"] = buf.actions.tensor.detach().numpy()
    out["terminals"] = buf.terminals.tensor.numpy()
    out["final_state"] = env.state.detach().numpy()
    for i, (kind, d) in enumerate(DRAWS[n_init:]):
        out[f"draw{i:03d}_{kind}"] = d.numpy()
    out["n_draws"] = np.array(len(DRAWS) - n_init)
    np.savez_compressed(HERE / "shac_cartpole_unsafe.npz", **out)
    print("wrote shac golden: loss", float(loss), "draws", len(DRAWS) - n_init, "terminals", buf.terminals.tensor.sum().item())
    # critic path on the SAME buffer (shac.py:154-251): TD-lambda targets, the critic fit, the Polyak update
    crit = {"td_weight": np.array(algo.td_weight), "vf_num_fits": np.array(algo.vf_num_fits),
            "vf_fit_num_batches": np.array(algo.vf_fit_num_batches), "polyak_target": np.array(algo.polyak_target)}
    with torch.no_grad():
        est, obs_rows = algo.calculate_estimated_values()
    crit["estimated_values"], crit["estimated_observations"] = est.numpy().copy(), obs_rows.numpy().copy()
    for k, v in algo.value_function.state_dict().items():
        crit[f"vf0.{k}"] = v.numpy().copy()
    crit["value_loss"] = np.array(algo.update_value_function())
    for k, v in algo.value_function.state_dict().items():
        crit[f"vf1.{k}"] = v.numpy().copy()
    algo.update_target_value_function()
    for k, v in algo.target_value_function.state_dict().items():
        crit[f"target_vf1.{k}"] = v.numpy().copy()
    np.savez_compressed(HERE / "shac_critic.npz", **crit)
    print("wrote critic golden: value loss", float(crit["value_loss"]), "targets", est.shape)


def buffer_golden():
    """tests/rollout_buffer_test.py:12-29 style known answers, produced by the reference's CoupledTensor."""
    from src.learning_algorithms.components.coupled_tensor import CoupledTensor
    ct = CoupledTensor(4, 3, 2)
    ct.tensor = torch.arange(24, dtype=torch.float64).reshape(4, 3, 2)
    idx = torch.tensor([0, 2, 3])
    got = ct[idx].numpy().copy()
    ct[idx] = torch.full((3, 2), -1.0)
    np.savez_compressed(HERE / "coupled_tensor.npz", idx=idx.numpy(), got=got, after=ct.tensor.numpy())


if __name__ == "__main__":
    from envs.balance_pendulum import BalancePendulumEnv
    from envs.balance_cartpole import BalanceCartPoleEnv
    from envs.swingup_cartpole import SwingUpCartPoleEnv
    from envs.balance_quadrotor import BalanceQuadrotorEnv
    from envs.manage_household import ManageHouseholdEnv
    for cls, name in [(BalancePendulumEnv, "BalancePendulum"), (BalanceCartPoleEnv, "BalanceCartPole"),
                      (SwingUpCartPoleEnv, "SwingUpCartPole"), (BalanceQuadrotorEnv, "BalanceQuadrotor"),
                      (ManageHouseholdEnv, "ManageHousehold")]:
        env_golden(cls, name, seed=3)
    shac_golden()
    buffer_golden()
