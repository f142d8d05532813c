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


if __name__ == "__main__" and "--timeline" not in sys.argv and "--gpu-timeline" not in sys.argv:
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


def gpu_timeline():
    """Device timeline of back-to-back graph-replayed episodes (torch.profiler / CUPTI): per kernel start, duration and the idle
    gap before it.  Shows where an episode's device time goes besides its kernels (launch gaps inside the graphs, the gap
    between the graphs, small copies)."""
    import bench
    import os
    from torch.profiler import profile, ProfilerActivity
    world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:          # torchrun: the same trace with the gradient all-reduce between the episode graph and the optimizer
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_default_dtype(torch.float32)
    torch.set_default_device(torch.device("cuda", local))
    torch.manual_seed(local)
    from safegbpo_b200.learning_algorithms.shac import SHAC
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.envs import balance_cartpole
    env = BoundaryProjectionSafeguard(balance_cartpole.BalanceCartPoleEnv(num_envs=4096, num_steps=1000000), strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": bench.POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": bench.VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=64, fused="on", rng_order="batched")
    if world > 1:
        from safegbpo_b200 import distributed as sgdist
        sgdist.attach(algo)

    def episode():
        algo.buffer.reset()
        algo.policy_optim.zero_grad()
        algo.collect_trajectories()
        return algo.update_policy()

    for _ in range(8):
        episode()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            episode()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, prev_end = evs[0].time_range.start, None
    rows = []
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = 0.0 if prev_end is None else s - prev_end
        rows.append((s, d, gap, e.name[:70]))
        prev_end = max(prev_end or 0.0, s + d)
    # the third episode: between the 3rd and 4th forward kernels
    starts = [i for i, r in enumerate(rows) if "rollout_wm" in r[3] or "rollout_fwd_kernel" in r[3] or "rollout_tc_fwd" in r[3]]
    a, b = starts[2], starts[3]
    first = a
    while first > 0 and rows[first - 1][0] > rows[starts[1]][0] and "adam" not in rows[first - 1][3]:
        first -= 1
    ep = rows[first:b]
    busy, gaps = sum(r[1] for r in ep), sum(r[2] for r in ep[1:])
    if local != 0:
        return
    print(f"episode span {ep[-1][0] + ep[-1][1] - ep[0][0]:.1f} us, kernels {busy:.1f} us, gaps {gaps:.1f} us, {len(ep)} device activities")
    for s, d, gap, name in ep:
        print(f"  +{s - ep[0][0]:8.1f} us  dur {d:7.1f}  gap before {gap:6.1f}  {name}")


if __name__ == "__main__" and "--gpu-timeline" in sys.argv:
    gpu_timeline()
