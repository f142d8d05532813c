"""GPU: the whole-horizon fused rollout (csrc/sgbpo_rollout*.cu) against (a) the reference's golden SHAC
episode, (b) the eager per-step path fed the same replayed RNG stream, in float64 (tight) and float32."""
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
        assert k == kind and tuple(d.shape) == tuple(shape), (k, kind, tuple(d.shape), shape)
        return d.to(self.dev) if isinstance(d, torch.Tensor) else torch.from_numpy(np.asarray(d)).to(self.dev)


class Recorder:
    def __init__(self):
        self.draws = []

    def __call__(self, kind, shape):
        t = torch.rand(*shape) if kind == "rand" else torch.randn(*shape)
        self.draws.append((kind, t.detach().clone()))
        return t


@pytest.fixture
def cuda_default():
    prev = torch.get_default_dtype()
    torch.set_default_device("cuda:0")
    yield torch.device("cuda:0")
    torch.set_default_device("cpu")
    torch.set_default_dtype(prev)
    from safegbpo_b200 import rng
    rng.set_source(None)


def test_fused_episode_vs_reference_golden(cuda_default, golden_dir):
    from safegbpo_b200 import rng
    from safegbpo_b200.envs.balance_cartpole import BalanceCartPoleEnv
    from safegbpo_b200.learning_algorithms.shac import SHAC
    torch.set_default_dtype(torch.float64)
    dev = cuda_default
    g = np.load(golden_dir / "shac_cartpole_unsafe.npz")
    keys = sorted(k for k in g.files if k.startswith("draw"))
    draws = [(k.split("_")[1], g[k]) for k in keys]
    env = BalanceCartPoleEnv(num_envs=4, num_steps=3)
    algo = SHAC(env=env, policy_kwargs={"net_arch": [32, 32]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [32, 32, 32]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=6, fused="on")
    algo.policy.load_state_dict({k[len("policy."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("policy.")})
    algo.target_value_function.load_state_dict(
        {k[len("target_vf."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("target_vf.")})
    env.state = torch.from_numpy(g["state0"]).to(dev)
    o0 = env.observation
    algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
    rng.set_source(Replay(draws, dev))
    algo.buffer.reset()
    algo.policy_optim.zero_grad()
    avg_reward = algo.collect_trajectories()
    assert algo._fused_engine is not None
    loss = algo._fused_engine.backward()
    tol = dict(rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(algo.buffer.observations.tensor.cpu().numpy(), g["observations"], **tol)
    np.testing.assert_allclose(algo.buffer.rewards.tensor.cpu().numpy(), g["rewards"], **tol)
    np.testing.assert_allclose(algo.buffer.values.tensor.cpu().numpy(), g["values"], **tol)
    np.testing.assert_allclose(algo.buffer.actions.tensor.cpu().numpy(), g["actions"], **tol)
    assert np.array_equal(algo.buffer.terminals.tensor.cpu().numpy(), g["terminals"])
    np.testing.assert_allclose(env.state.cpu().numpy(), g["final_state"], **tol)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-11)
    np.testing.assert_allclose(avg_reward, float(g["avg_reward"]), rtol=1e-11)
    for k, p in algo.policy.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), g[f"grad.{k}"], rtol=1e-8, atol=1e-12, err_msg=k)


CONFIGS = [
    ("BalanceCartPole", "bp", {"gate": "reference"}, 64, 5),
    ("BalanceCartPole", "bp", {"gate": "always"}, 64, 1000),
    ("SwingUpCartPole", "rm", {"gate": "always", "zono": True, "linear": True}, 64, 1000),
    ("SwingUpCartPole", "rm", {"gate": "always", "zono": False, "linear": False}, 32, 1000),
    ("BalancePendulum", "none", {}, 64, 7),
    ("BalancePendulum", "rm", {"gate": "reference", "zono": True, "linear": True}, 64, 1000),     # BASELINE configs[0]
    ("BalancePendulum", "rm", {"gate": "always", "zono": True, "linear": True}, 32, 1000),
    ("BalancePendulum", "bp", {"gate": "always"}, 32, 9),
    ("BalanceQuadrotor", "none", {}, 32, 6),
]


def build(env_name, sg, opts, width, num_steps, N, H, fused, dev):
    from safegbpo_b200.envs.balance_cartpole import BalanceCartPoleEnv
    from safegbpo_b200.envs.swingup_cartpole import SwingUpCartPoleEnv
    from safegbpo_b200.envs.balance_pendulum import BalancePendulumEnv
    from safegbpo_b200.envs.balance_quadrotor import BalanceQuadrotorEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    from safegbpo_b200.learning_algorithms.shac import SHAC
    cls = {"BalanceCartPole": BalanceCartPoleEnv, "SwingUpCartPole": SwingUpCartPoleEnv,
           "BalancePendulum": BalancePendulumEnv, "BalanceQuadrotor": BalanceQuadrotorEnv}[env_name]
    env = cls(num_envs=N, num_steps=num_steps)
    wrapped = env
    if sg == "bp":
        wrapped = BoundaryProjectionSafeguard(env, gate=opts["gate"], strict=False)
    elif sg == "rm":
        wrapped = RayMaskSafeguard(env, linear_projection=opts["linear"], zonotopic_approximation=opts["zono"],
                                   passthrough=False, gate=opts["gate"], strict=False)
    algo = SHAC(env=wrapped, policy_kwargs={"net_arch": [width, width]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [width, width, width]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=H, fused=fused)
    return env, wrapped, algo


@pytest.mark.parametrize("env_name,sg,opts,width,num_steps", CONFIGS,
                         ids=[f"{c[0]}-{c[1]}-w{c[3]}-T{c[4]}" for c in CONFIGS])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_fused_equals_eager(cuda_default, env_name, sg, opts, width, num_steps, dtype):
    """Same weights, same initial states, same replayed draws: the fused kernels and the eager per-step path
    must produce the same buffers, loss and policy gradients (resets inside the horizon included)."""
    from safegbpo_b200 import rng
    torch.set_default_dtype(dtype)
    dev = cuda_default
    N, H = 40, 12        # N not a multiple of 8 x 4: exercises the ragged last warp / CTA
    torch.manual_seed(3)
    env_e, wrap_e, algo_e = build(env_name, sg, opts, width, num_steps, N, H, "off", dev)
    env_f, wrap_f, algo_f = build(env_name, sg, opts, width, num_steps, N, H, "on", dev)
    algo_f.policy.load_state_dict(algo_e.policy.state_dict())
    algo_f.target_value_function.load_state_dict(algo_e.target_value_function.state_dict())
    for wrap, env in ((wrap_e, env_e), (wrap_f, env_f)):
        if wrap is not env:
            # keep the reference's failure value (NaN) for infeasible programs: both paths must then agree on WHICH environments
            # fail; the non-strict default (raw action executed, rollout continues) would let one borderline float32 verdict
            # send the two trajectories apart
            wrap.consts().infeasible_raw = 0
    # start inside the feasible box (RCI resets put cartpole far outside; keep programs feasible for the comparison)
    half = env_e.state_set.generator[0].diagonal()
    spread = 0.1 if env_name == "BalancePendulum" else 0.6      # the pendulum's RCI set is a small part of the box
    s0 = env_e.state_set.center[0] + (torch.rand(N, env_e.state_dim) * 2 - 1) * half * spread
    for env, algo in ((env_e, algo_e), (env_f, algo_f)):
        env.state, env.steps, env._reset_pending = s0.clone(), 2, False
        if hasattr(env, "goal"):
            env.goal = s0[:, :2].clone()
        o0 = env.observation
        algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
    rec = Recorder()
    rng.set_source(rec)
    algo_e.buffer.reset()
    algo_e.policy_optim.zero_grad()
    r_e = algo_e.collect_trajectories()
    loss_e = algo_e.policy_loss()
    loss_e.backward()
    rng.set_source(Replay(rec.draws, dev))
    algo_f.buffer.reset()
    algo_f.policy_optim.zero_grad()
    r_f = algo_f.collect_trajectories()
    loss_f = algo_f._fused_engine.backward()
    torch.cuda.synchronize()
    if dtype == torch.float64:
        tol, gtol = dict(rtol=1e-9, atol=1e-10), dict(rtol=1e-7, atol=1e-9)
    else:
        tol, gtol = dict(rtol=2e-4, atol=2e-4), dict(rtol=5e-3, atol=5e-4)      # fp32 over 12 chained steps
        if sg == "rm":
            # the ray-mask maps divide by a chord length; fp32 rounding differences between torch's MLP and the
            # fused tiles are amplified by that factor at every one of the 12 chained steps (the f64 run of this
            # same case pins the logic to 1e-9, test_gpu_step_parity pins the single step in fp32)
            tol, gtol = dict(rtol=1e-2, atol=2e-3), dict(rtol=5e-2, atol=5e-3)
    be, bf = algo_e.buffer, algo_f.buffer
    assert torch.equal(be.terminals.tensor, bf.terminals.tensor)
    fin_e = torch.isfinite(be.observations.tensor).all(dim=2).all(dim=0)
    fin_f = torch.isfinite(bf.observations.tensor).all(dim=2).all(dim=0)
    finite = fin_e & fin_f
    if sg == "rm" and opts.get("gate") == "always":
        # the ray mask AS SHIPPED (Q2) can push the state out of the RCI set, after which the programs are
        # infeasible (NaN rows, FL_INFEASIBLE); both paths must agree on which environments that happens to
        assert (fin_e == fin_f).float().mean() > 0.95 and finite.any()
    else:
        assert finite.float().mean() > 0.9
    np.testing.assert_allclose(bf.observations.tensor[:, finite].cpu().numpy(), be.observations.tensor[:, finite].detach().cpu().numpy(), **tol)
    np.testing.assert_allclose(bf.actions.tensor[:, finite].cpu().numpy(), be.actions.tensor[:, finite].detach().cpu().numpy(), **tol)
    np.testing.assert_allclose(bf.rewards.tensor[:, finite].cpu().numpy(), be.rewards.tensor[:, finite].detach().cpu().numpy(), **tol)
    np.testing.assert_allclose(bf.values.tensor[:, finite].cpu().numpy(), be.values.tensor[:, finite].detach().cpu().numpy(), **tol)
    if bool(finite.all()):
        np.testing.assert_allclose(float(loss_f), float(loss_e), rtol=tol["rtol"])
        np.testing.assert_allclose(r_f, r_e, rtol=tol["rtol"])
        for (k, pe), (_, pf) in zip(algo_e.policy.named_parameters(), algo_f.policy.named_parameters()):
            scale = float(pe.grad.abs().max()) + 1e-30
            np.testing.assert_allclose(pf.grad.cpu().numpy() / scale, pe.grad.cpu().numpy() / scale, err_msg=k,
                                       rtol=gtol["rtol"], atol=gtol["atol"])
        if wrap_e is not env_e:
            assert wrap_e.interventions == wrap_f.interventions


def test_mlp_rows_matches_torch(cuda_default):
    import ctypes as C
    from safegbpo_b200 import _lib
    from safegbpo_b200.learning_algorithms.components.value_function import ValueFunction
    from safegbpo_b200.learning_algorithms.fused_rollout import PackedMlp, mlp_rows_fits
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 2e-5)):
        torch.set_default_dtype(dtype)
        for width, depth, M in ((32, 1, 5), (64, 2, 1000), (128, 3, 4099)):
            if not mlp_rows_fits(dtype, 5, width, depth):
                continue
            vf = ValueFunction(5, net_arch=[width] * depth)
            x = torch.randn(M, 5) * 3
            ref = vf(x).squeeze(1)
            pack = PackedMlp(vf.model)
            out = torch.empty(M)
            _lib.check(_lib.lib().sgbpo_mlp_rows(_lib.dtype_code(dtype), M, C.byref(pack.desc), C.c_void_p(x.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), _lib.stream_ptr()), "sgbpo_mlp_rows")
            torch.cuda.synchronize()
            np.testing.assert_allclose(out.cpu().numpy(), ref.detach().cpu().numpy(), rtol=tol, atol=tol)


@pytest.mark.parametrize("env_name", ["BalanceCartPole", "BalanceQuadrotor"])
def test_graph_replay_with_resets_inside_the_horizon(cuda_default, env_name):
    """Episodes whose horizon contains reset-all events (Q6; num_steps = 10 < H = 16, so every episode has one or
    two) replay the SAME graphs: the schedule (reset / aux indices, terminal rows, discounts) is uploaded per
    episode and the reset states are drawn inside the graph.  Must equal the eager launches of the batched RNG mode."""
    torch.set_default_dtype(torch.float32)
    dev = cuda_default
    N, H = 96, 16
    results = []
    for use_graphs in (False, True):
        torch.manual_seed(21)
        sg = "bp" if env_name == "BalanceCartPole" else "none"
        env, wrapped, algo = build(env_name, sg, {"gate": "reference"}, 64, 10, N, H, "on", dev)
        algo.rng_order = "batched"
        half = env.state_set.generator[0].diagonal()
        env.state = env.state_set.center[0] + (torch.rand(N, env.state_dim) * 2 - 1) * half * 0.3
        if hasattr(env, "goal"):
            env.goal = env.state[:, :2].clone()
        env.steps, env._reset_pending = 0, False
        o0 = env.observation
        algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
        torch.manual_seed(6)
        losses, rewards, used = [], [], 0
        for ep in range(6):
            algo.buffer.reset()
            algo.policy_optim.zero_grad()
            rewards.append(algo.collect_trajectories())
            eng = algo._fused_engine
            eng.use_graphs, eng.graph_warmup = use_graphs, 1
            used += bool(eng.last.get("graph"))
            losses.append(algo.update_policy())
        results.append((losses, rewards, algo.buffer.observations.tensor.clone(), algo.buffer.terminals.tensor.clone(),
                        [p.detach().clone() for p in algo.policy.parameters()], env.steps, env._reset_pending, used,
                        env.goal.clone() if hasattr(env, "goal") else None))
    (l0, r0, o0_, t0, p0, s0, pend0, u0, goal0), (l1, r1, o1_, t1, p1, s1, pend1, u1, goal1) = results
    assert u0 == 0 and u1 >= 4              # graph replays really happened, with resets inside
    assert (s0, pend0) == (s1, pend1) and torch.equal(t0, t1) and t0[1:H].any()
    np.testing.assert_allclose(l1, l0, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(r1, r0, rtol=2e-4, atol=1e-6)
    # (BalanceQuadrotor: the reward's sqrt has an infinite slope at the goal = reset position, so the reference's
    #  own gradients go non-finite there; both paths must agree on that too -> equal_nan)
    assert torch.allclose(o0_, o1_, rtol=2e-3, atol=2e-3, equal_nan=True)
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-5, equal_nan=True)
    if goal0 is not None:
        assert torch.allclose(goal0, goal1, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rng_order", ["batched", "reference"])
def test_graph_replay_equals_eager_launches(cuda_default, rng_order):
    """Four consecutive training episodes (collect + update_policy incl. Adam): the CUDA-graph replay of the
    fused engine must give bit-identical results to launching the same kernels eagerly."""
    torch.set_default_dtype(torch.float32)
    dev = cuda_default
    N, H = 256, 16
    results = []
    for use_graphs in (False, True):
        torch.manual_seed(11)
        env, wrapped, algo = build("BalanceCartPole", "bp", {"gate": "reference"}, 64, 1000, N, H, "on", dev)
        algo.rng_order = rng_order
        half = env.state_set.generator[0].diagonal()
        env.state = env.state_set.center[0] + (torch.rand(N, 4) * 2 - 1) * half * 0.5
        env.steps, env._reset_pending = 0, False
        o0 = env.observation
        algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
        torch.manual_seed(5)
        losses = []
        for ep in range(5):
            algo.buffer.reset()
            algo.policy_optim.zero_grad()
            algo.collect_trajectories()
            eng = algo._fused_engine
            eng.use_graphs, eng.graph_warmup = use_graphs, 1
            losses.append(algo.update_policy())
        results.append((losses, algo.buffer.observations.tensor.clone(), [p.detach().clone() for p in algo.policy.parameters()],
                        wrapped.interventions, eng._g is not None and eng._g["fwd_graph"] is not None))
    (l0, o0_, p0, i0, g0), (l1, o1_, p1, i1, g1) = results
    assert not g0 and g1
    np.testing.assert_allclose(l1, l0, rtol=1e-6)
    # identical kernels; the only arithmetic difference is how the noise sample is assembled from the uniform
    # draw (addcmul in the captured body), i.e. 1-ulp differences amplified over 5 x 16 chained steps
    assert torch.allclose(o0_, o1_, rtol=1e-3, atol=1e-3)
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)
    assert i0 == i1


def test_full_size_rollout_is_consistent_with_the_step_kernels(cuda_default):
    """BASELINE configs[1] at full size (BalanceCartPole + boundary projection, 4096 environments, H = 64, fp32):
    size-independent properties of the fused whole-horizon kernels -- every transition of the rollout buffer is
    reproduced by the per-step kernel from (s_t, a_t, w_t) (same executed action, next state, observation, reward,
    flag word), the chain is closed (row t + 1 of the state buffer is what step t produced), no feasible guarded
    action leaves its sets, and the reverse pass is repeatable."""
    from safegbpo_b200 import _lib, ops
    torch.set_default_dtype(torch.float32)
    dev = cuda_default
    N, H = 4096, 64
    torch.manual_seed(17)
    env, wrapped, algo = build("BalanceCartPole", "bp", {"gate": "reference"}, 128, 240, N, H, "on", dev)
    algo.rng_order = "batched"
    env.reset(seed=3)
    env.steps, env._reset_pending = 200, False                 # a reset-all falls inside the horizon (t = 40)
    o0 = env.observation
    algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
    eng = algo._engine()
    eng.use_graphs = False
    algo.buffer.reset()
    algo.policy_optim.zero_grad()
    algo.collect_trajectories()
    last = eng.last
    states, obs, rewards, flags, a_exec = last["states"], last["obs"], last["rewards"], last["flags"], last["exec_actions"]
    actions = algo.buffer.actions.tensor
    eps, noise = last["keep"][-9], last["keep"][-8]
    assert tuple(noise.shape) == (H, N, 4) and tuple(states.shape) == (H + 1, N, 4)
    consts = wrapped.consts()
    checked = 0
    for t in (0, 1, 17, 38, 41, 63):
        s_next, ae, ob, rw, fl = ops.safeguarded_step(states[t], actions[t], noise[t], env.env_id, consts)
        is_reset_step = not torch.allclose(states[t + 1], s_next, rtol=1e-5, atol=1e-5)
        assert torch.equal(fl, flags[t]), f"flag words differ at t={t}"
        np.testing.assert_allclose(ae.cpu().numpy(), a_exec[t].cpu().numpy(), rtol=1e-6, atol=1e-7)
        if not is_reset_step:                                  # (on the reset step the buffer holds the re-sampled state)
            np.testing.assert_allclose(s_next.cpu().numpy(), states[t + 1].cpu().numpy(), rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ob.cpu().numpy(), obs[t + 1].cpu().numpy(), rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(rw.cpu().numpy(), rewards[t].cpu().numpy(), rtol=1e-5, atol=1e-5)
            checked += 1
    assert checked >= 5
    guard = (flags & _lib.FL_GUARD_USED) != 0
    feas = (flags & _lib.FL_INFEASIBLE) == 0
    good = guard & feas
    assert int(good.sum()) > 1000
    assert bool((((flags & _lib.FL_NEXT_IN_SET) != 0) | ~good).all())          # zero safety violations where feasible
    assert bool((((flags & _lib.FL_ACTION_IN_SET) != 0) | ~good).all())
    # the reverse pass is a pure function of the buffers and the stash: running it twice gives the same gradients
    loss1 = eng.backward()
    g1 = [p.grad.clone() for p in algo.policy.parameters()]
    for p in algo.policy.parameters():
        p.grad = None
    loss2 = eng.backward()
    g2 = [p.grad.clone() for p in algo.policy.parameters()]
    for a, b in zip(g1, g2):                                   # deterministic up to the atomic accumulation order
        assert torch.allclose(a, b, rtol=2e-3, atol=1e-6 * float(a.abs().max() + 1e-30) + 1e-9)
    assert float(loss1) == float(loss2)
