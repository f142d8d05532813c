"""Tuning aid: where one graph-replayed SHAC episode spends its time (CUDA events around the two graph
replays and around the whole episode)."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402


def main():
    import bench
    from safegbpo_b200.learning_algorithms.shac import SHAC
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.envs import balance_cartpole
    dev = torch.device("cuda", 0)
    torch.set_default_dtype(torch.float32)
    torch.set_default_device(dev)
    torch.manual_seed(0)
    env = BoundaryProjectionSafeguard(balance_cartpole.BalanceCartPoleEnv(num_envs=4096, num_steps=1000000), strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": bench.POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": bench.VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=64, fused="on", rng_order="batched")

    def episode():
        algo.buffer.reset()
        algo.policy_optim.zero_grad()
        algo.collect_trajectories()
        return algo.update_policy()

    for _ in range(6):
        episode()
    torch.cuda.synchronize()
    g = algo._fused_engine._g
    assert g is not None and g["fwd_graph"] is not None and g["bwd_graph"] is not None

    def ev_time(fn, n=20):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    out = {"fwd_graph_ms": ev_time(g["fwd_graph"].replay), "bwd_graph_ms": ev_time(g["bwd_graph"].replay)}

    def tail():
        torch.nn.utils.clip_grad_norm_(algo.policy.parameters(), 1.0)
        algo.policy_optim.step()
    out["clip_adam_ms"] = ev_time(tail)
    out["episode_ms"] = ev_time(episode)
    print(json.dumps(out))


if __name__ == "__main__" and "--timeline" not in sys.argv:
    main()


def host_timeline():
    """Host-side wall clock of each stage of one graph-replayed episode (averaged)."""
    import bench
    torch.set_default_dtype(torch.float32)
    torch.set_default_device(torch.device("cuda", 0))
    torch.manual_seed(0)
    from safegbpo_b200.learning_algorithms.shac import SHAC
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.envs import balance_cartpole
    env = BoundaryProjectionSafeguard(balance_cartpole.BalanceCartPoleEnv(num_envs=4096, num_steps=1000000), strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": bench.POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": bench.VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=64, fused="on", rng_order="batched")
    eng = algo._engine()
    marks = {}

    def stamp(name, t0):
        marks.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    import safegbpo_b200.optim as sgoptim
    for it in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        algo.buffer.reset(); stamp("buffer.reset", t0)
        algo.policy_optim.zero_grad(); stamp("zero_grad", t0)
        algo.collect_trajectories(); stamp("collect(returns after reward sync)", t0)
        loss = eng.backward(); stamp("engine.backward (attach grads)", t0)
        sgoptim.clip_adam_step(algo.policy_optim, 1.0); stamp("clip_adam launch", t0)
        float(loss); stamp("float(loss) sync", t0)
    print(json.dumps({k: round(sum(v[10:]) / len(v[10:]), 4) for k, v in marks.items()}))


if __name__ == "__main__" and "--timeline" in sys.argv:
    host_timeline()
