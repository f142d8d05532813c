"""GPU: the tensor-core rollout engine (csrc/sgbpo_rollout_tc.cuh: tcgen05 forward, parallel step Jacobians, tcgen05
reverse pass) against the float64 evaluation of the same episode.

Method (self-calibrating, no hand-picked chaos allowance): the SAME episode -- weights, initial states, replayed
draws -- runs three times through the fused path: float64 on the FMA tiles (the truth; pinned to the reference's golden
episode and to the eager autograd path elsewhere in this suite), float32 on the FMA tiles (engine 0) and float32 on the
tensor cores (engine 1).  The float32 FMA run measures how much float32 rounding the chained, unstable rollout
amplifies; the tensor-core run must stay within a small multiple of it (3x) or under the stated absolute bar
(1e-5 relative per step -> 2e-4 over the chained horizon), whichever is larger.
"""
import os

import numpy as np
import pytest
import torch

from test_gpu_fused_rollout import Recorder, Replay, build, cuda_default  # noqa: F401



def _episode(env_name, sg, opts, N, H, num_steps, dtype, engine, dev, state, weights, draws, spread, r_rows=None):
    """One collect + backward through the fused path; returns buffers, loss, gradients."""
    from safegbpo_b200 import rng
    torch.set_default_dtype(dtype)
    os.environ["SGBPO_ROLLOUT_ENGINE"] = str(engine)
    if r_rows:
        os.environ["SGBPO_ROLLOUT_R"] = str(r_rows)
    try:
        torch.manual_seed(3)
        env, wrapped, algo = build(env_name, sg, opts, 128, num_steps, N, H, "on", dev)
        if weights is None:
            weights = ({k: v.double() for k, v in algo.policy.state_dict().items()},
                       {k: v.double() for k, v in algo.target_value_function.state_dict().items()})
        algo.policy.load_state_dict({k: v.to(dtype) for k, v in weights[0].items()})
        algo.target_value_function.load_state_dict({k: v.to(dtype) for k, v in weights[1].items()})
        if state is None:
            half = env.state_set.generator[0].diagonal().double()
            state = env.state_set.center[0].double() + (torch.rand(N, env.state_dim, dtype=torch.float64) * 2 - 1) * half * spread
        env.state, env.steps, env._reset_pending = state.to(dtype).clone(), 2, False
        if hasattr(env, "goal"):
            env.goal = state[:, :2].to(dtype).clone()
        o0 = env.observation
        algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
        if draws is None:
            rec = Recorder()
            rng.set_source(rec)
        else:
            rng.set_source(Replay([(k, d.to(dtype)) for k, d in draws], dev))
        algo.buffer.reset()
        algo.policy_optim.zero_grad()
        algo.collect_trajectories()
        eng = algo._fused_engine
        want = {0: (0, 0), 1: (1, 1), 2: (0, 1), 3: (2, 1)}[engine] if dtype == torch.float32 else (0, 0)
        assert eng is not None and (eng.engine_fwd, eng.engine_bwd) == want
        loss = eng.backward()
        torch.cuda.synchronize()
        if draws is None:
            draws = [(k, d.double()) for k, d in rec.draws]
        out = dict(obs=algo.buffer.observations.tensor.double().cpu(), actions=algo.buffer.actions.tensor.double().cpu(),
                   rewards=algo.buffer.rewards.tensor.double().cpu(), loss=float(loss), flags=eng.last["flags"].cpu(),
                   grads={k: p.grad.double().cpu() for k, p in algo.policy.named_parameters()})
        return out, state, weights, draws
    finally:
        os.environ.pop("SGBPO_ROLLOUT_ENGINE", None)
        os.environ.pop("SGBPO_ROLLOUT_R", None)
        rng.set_source(None)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-300))


CASES = [
    # env, safeguard, options, N, H, num_steps, spread of the initial states inside the feasible box
    ("BalanceCartPole", "bp", {"gate": "reference"}, 300, 12, 5, 0.6),            # ragged last tile, resets inside the horizon
    ("BalanceCartPole", "bp", {"gate": "always"}, 256, 12, 1000, 0.6),
    ("SwingUpCartPole", "rm", {"gate": "always", "zono": False, "linear": False}, 130, 12, 1000, 0.6),
    ("BalancePendulum", "rm", {"gate": "reference", "zono": True, "linear": True}, 64, 12, 1000, 0.1),   # BASELINE configs[0]
    ("BalancePendulum", "none", {}, 200, 12, 7, 0.1),
    ("BalanceQuadrotor", "none", {}, 140, 12, 6, 0.3),
]


@pytest.mark.gpu
@pytest.mark.parametrize("env_name,sg,opts,N,H,num_steps,spread", CASES, ids=[f"{c[0]}-{c[1]}-N{c[3]}" for c in CASES])
def test_tensor_core_engine_vs_float64(cuda_default, env_name, sg, opts, N, H, num_steps, spread):
    dev = cuda_default
    ref, state, weights, draws = _episode(env_name, sg, opts, N, H, num_steps, torch.float64, 0, dev, None, None, None, spread)
    fma, *_ = _episode(env_name, sg, opts, N, H, num_steps, torch.float32, 0, dev, state, weights, draws, spread)
    for engine in (1, 2, 3):     # tcgen05 forward / FMA forward / warp-MMA forward, each + parallel reverse pass
        tc, *_ = _episode(env_name, sg, opts, N, H, num_steps, torch.float32, engine, dev, state, weights, draws, spread)
        fin = torch.isfinite(ref["obs"]).all(dim=2).all(dim=0) & torch.isfinite(fma["obs"]).all(dim=2).all(dim=0) \
            & torch.isfinite(tc["obs"]).all(dim=2).all(dim=0)
        assert fin.float().mean() > 0.9
        for key in ("obs", "actions", "rewards"):
            r = ref[key][:, fin]
            e_fma, e_tc = _rel(fma[key][:, fin], r), _rel(tc[key][:, fin], r)
            assert e_tc <= max(3.0 * e_fma, 2e-4), f"engine {engine} {key}: tensor-core {e_tc:.2e} vs FMA {e_fma:.2e}"
        if bool(fin.all()) and all(bool(torch.isfinite(g).all()) for g in ref["grads"].values()):
            assert abs(tc["loss"] - ref["loss"]) <= max(3.0 * abs(fma["loss"] - ref["loss"]), 2e-4 * abs(ref["loss"]))
            for k, g in ref["grads"].items():
                e_fma, e_tc = _rel(fma["grads"][k], g), _rel(tc["grads"][k], g)
                assert e_tc <= max(3.0 * e_fma, 5e-4), f"engine {engine} grad {k}: tensor-core {e_tc:.2e} vs FMA {e_fma:.2e}"


@pytest.mark.gpu
@pytest.mark.parametrize("engine,r_rows", [(1, None), (2, 4), (3, None), (0, 4), (0, 8)],
                         ids=["tcgen05", "fma-fwd+tc-reverse", "warp-mma-fwd+tc-reverse", "fma-R4", "fma-R8"])
def test_headline_instantiations_gradients_vs_float64(cuda_default, engine, r_rows):
    """BASELINE configs[1] shapes (BalanceCartPole + boundary projection, policy [128, 128], H = 64, N = 1024): the
    policy gradients of the instantiations the benchmark times -- rollout_tc_bwd_kernel and
    rollout_bwd_kernel<1, float, 128, R in {4, 8}> -- against the float64 episode on the same draws."""
    dev = cuda_default
    args = ("BalanceCartPole", "bp", {"gate": "reference"}, 1024, 64, 240)
    ref, state, weights, draws = _episode(*args, torch.float64, 0, dev, None, None, None, 0.3)
    got, *_ = _episode(*args, torch.float32, engine, dev, state, weights, draws, 0.3, r_rows=r_rows)
    fin = torch.isfinite(ref["obs"]).all(dim=2).all(dim=0) & torch.isfinite(got["obs"]).all(dim=2).all(dim=0)
    assert bool(fin.all())
    # float32 rounding amplified over 64 chained steps of an unstable plant: trajectories agree to ~1e-3, the episode's
    # summed gradients to a few 1e-3 of their norm (measured: see profiles/r2_parity.md)
    assert _rel(got["obs"], ref["obs"]) < 5e-3
    assert abs(got["loss"] - ref["loss"]) < 2e-3 * abs(ref["loss"])
    for k, g in ref["grads"].items():
        assert _rel(got["grads"][k], g) < 2e-2, f"grad {k}: {_rel(got['grads'][k], g):.2e}"


def test_tensor_core_kernels_are_tcgen05():
    """The rollout objects contain the Blackwell tensor-core path: UTCHMMA (tcgen05.mma) and LDTM (tcgen05.ld)."""
    import subprocess
    from safegbpo_b200 import _lib
    obj = _lib.PKG / "build" / "sgbpo_rollout_tc_env1.o"
    if not obj.exists():
        pytest.skip("object files are not shipped to the GPU box (the CPU run checks them)")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    for kernel in ("rollout_tc_fwd_kernel", "policy_rows_bwd_kernel"):
        body = sass[sass.index(kernel):]
        body = body[:body.index("Function :", 10)] if "Function :" in body[10:] else body
        assert "UTCHMMA" in body and "LDTM" in body, kernel


def test_warp_mma_kernel_uses_tensor_core_mma():
    """Forward engine 2 issues warp-level tensor-core MMAs (HMMA.16816.F32) and has no block-wide barrier in its time loop."""
    import subprocess
    from safegbpo_b200 import _lib
    obj = _lib.PKG / "build" / "sgbpo_rollout_wm_env1.o"
    if not obj.exists():
        pytest.skip("object files are not shipped to the GPU box (the CPU run checks them)")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    funcs = {}
    for chunk in sass.split("Function : ")[1:]:
        funcs[chunk.split("\n", 1)[0].strip()] = chunk
    one = next(v for k, v in funcs.items() if "rollout_wm_fwd_kernel" in k)
    two = next(v for k, v in funcs.items() if "rollout_wm2_fwd_kernel" in k)
    assert one.count("HMMA.16816.F32") >= 288 and one.count("BAR.SYNC") <= 2      # the parameter staging barrier only
    assert two.count("HMMA.16816.F32") >= 144 and "BAR.SYNC" in two               # pair-split: named 64-thread barriers


def _build_any(env_name, N, H, num_steps, fused, dev):
    from safegbpo_b200.envs.manage_household import ManageHouseholdEnv
    from safegbpo_b200.envs.navigate_seeker import NavigateSeekerEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.learning_algorithms.shac import SHAC
    if env_name == "ManageHousehold":
        env = ManageHouseholdEnv(num_envs=N, num_steps=num_steps)
    else:
        env = NavigateSeekerEnv(num_envs=N, num_steps=num_steps, num_obstacles=1, min_radius=2.0, max_radius=4.0)
    wrapped = BoundaryProjectionSafeguard(env, gate="always", strict=False)
    algo = SHAC(env=wrapped, policy_kwargs={"net_arch": [128, 128]}, policy_optim_kwargs={"lr": 1e-3},
                vf_kwargs={"net_arch": [128, 128, 128]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
                len_trajectories=H, fused=fused)
    return env, wrapped, algo


@pytest.mark.gpu
@pytest.mark.parametrize("env_name,num_steps", [("ManageHousehold", 720), ("NavigateSeeker", 400), ("NavigateSeeker", 5)],
                         ids=["household", "seeker", "seeker-resets"])
def test_household_and_seeker_fused_vs_eager(cuda_default, env_name, num_steps):
    """ManageHousehold (per-step shared series, 23 observations) and NavigateSeeker (obstacle columns of the observation,
    per-step safe action set P5, collision operator, goal + obstacles re-sampled at every reset) through the tensor-core
    whole-horizon engine against the eager per-step path (torch autograd over the per-step kernels) on replayed draws:
    float64 eager = truth, float32 eager = the float32 noise floor, float32 fused must stay within 3x of it."""
    from safegbpo_b200 import rng
    dev = cuda_default
    N, H = 200, 10
    out = {}
    draws = state = weights = aux = None
    for tag, dtype, fused in (("ref", torch.float64, "off"), ("eager32", torch.float32, "off"), ("fused32", torch.float32, "on")):
        torch.set_default_dtype(dtype)
        torch.manual_seed(5)
        env, wrapped, algo = _build_any(env_name, N, H, num_steps, fused, dev)
        if weights is None:
            weights = ({k: v.double() for k, v in algo.policy.state_dict().items()},
                       {k: v.double() for k, v in algo.target_value_function.state_dict().items()})
            half = env.state_set.generator[0].diagonal().double()
            state = env.state_set.center[0].double() + (torch.rand(N, env.state_dim, dtype=torch.float64) * 2 - 1) * half * 0.5
            if env_name == "NavigateSeeker":
                env.state = state.clone()
                env._after_reset()                  # goal + obstacles for these states (host rejection sampling)
                aux = (env.goal.double().clone(), [(b.center.double().clone(), b.radius.double().clone()) for b in env.obstacles])
        algo.policy.load_state_dict({k: v.to(dtype) for k, v in weights[0].items()})
        algo.target_value_function.load_state_dict({k: v.to(dtype) for k, v in weights[1].items()})
        env.state, env.steps, env._reset_pending = state.to(dtype).clone(), 2, False
        if aux is not None:
            env.goal = aux[0].to(dtype).clone()
            for b, (c, r) in zip(env.obstacles, aux[1]):
                b.center, b.radius = c.to(dtype).clone(), r.to(dtype).clone()
        o0 = env.observation
        algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
        if draws is None:
            rec = Recorder()
            rng.set_source(rec)
        else:
            rng.set_source(Replay([(k, d.to(dtype)) for k, d in draws], dev))
        algo.buffer.reset()
        algo.policy_optim.zero_grad()
        algo.collect_trajectories()
        if fused == "on":
            eng = algo._fused_engine
            assert eng is not None and (eng.engine_fwd, eng.engine_bwd) == (1, 1)
            loss = eng.backward()
        else:
            loss = algo.policy_loss()
            loss.backward()
        torch.cuda.synchronize()
        rng.set_source(None)
        if draws is None:
            draws = [(k, d.double()) for k, d in rec.draws]
        out[tag] = dict(obs=algo.buffer.observations.tensor.detach().double().cpu(), rewards=algo.buffer.rewards.tensor.detach().double().cpu(),
                        actions=algo.buffer.actions.tensor.detach().double().cpu(), loss=float(loss),
                        grads={k: p.grad.double().cpu() for k, p in algo.policy.named_parameters()})
    ref, e32, f32 = out["ref"], out["eager32"], out["fused32"]
    fin = torch.isfinite(ref["obs"]).all(dim=2).all(dim=0) & torch.isfinite(e32["obs"]).all(dim=2).all(dim=0) \
        & torch.isfinite(f32["obs"]).all(dim=2).all(dim=0)
    assert fin.float().mean() > 0.9
    for key in ("obs", "actions", "rewards"):
        r = ref[key][:, fin]
        a, b = _rel(e32[key][:, fin], r), _rel(f32[key][:, fin], r)
        assert b <= max(3.0 * a, 2e-4), f"{key}: fused {b:.2e} vs eager float32 {a:.2e}"
    if bool(fin.all()):
        assert abs(f32["loss"] - ref["loss"]) <= max(3.0 * abs(e32["loss"] - ref["loss"]), 2e-4 * abs(ref["loss"]))
        for k, g in ref["grads"].items():
            a, b = _rel(e32["grads"][k], g), _rel(f32["grads"][k], g)
            assert b <= max(3.0 * a, 1e-3), f"grad {k}: fused {b:.2e} vs eager float32 {a:.2e}"
