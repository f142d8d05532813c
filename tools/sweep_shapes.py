"""Tuning aid: kernel times of the fused rollout for forced (rows per warp, warps per CTA) launch shapes.
    python tools/sweep_shapes.py [--env BalanceCartPole] [--num-envs 4096] [--horizon 64]
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--env", default="BalanceCartPole")
    p.add_argument("--safeguard", default="boundary_projection")
    p.add_argument("--num-envs", type=int, default=4096)
    p.add_argument("--horizon", type=int, default=64)
    p.add_argument("--shapes", default="8x4,8x7,4x7,4x12,2x14,2x16")
    p.add_argument("--rows", default="8x8,4x16,4x12")
    a = p.parse_args()
    import bench
    from safegbpo_b200.learning_algorithms.shac import SHAC
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    from safegbpo_b200.envs import balance_cartpole, balance_pendulum, swingup_cartpole, balance_quadrotor
    dev = torch.device("cuda", 0)
    torch.set_default_dtype(torch.float32)
    torch.set_default_device(dev)
    torch.manual_seed(0)
    cls = {"BalanceCartPole": balance_cartpole.BalanceCartPoleEnv, "BalancePendulum": balance_pendulum.BalancePendulumEnv,
           "SwingUpCartPole": swingup_cartpole.SwingUpCartPoleEnv, "BalanceQuadrotor": balance_quadrotor.BalanceQuadrotorEnv}[a.env]
    env = cls(num_envs=a.num_envs, num_steps=100000)
    if a.safeguard == "boundary_projection":
        env = BoundaryProjectionSafeguard(env, strict=False)
    elif a.safeguard == "ray_mask":
        env = RayMaskSafeguard(env, linear_projection=True, zonotopic_approximation=True, passthrough=False, strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": bench.POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": bench.VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=a.horizon, fused="on", rng_order="batched")
    engine = algo._engine()
    engine.use_graphs = False

    def run(n):
        for _ in range(n):
            algo.buffer.reset()
            algo.policy_optim.zero_grad()
            algo.collect_trajectories()
            algo.update_policy()
        torch.cuda.synchronize()

    def measure(tag):
        run(2)
        engine.kernel_ms = {k: [] for k in engine.kernel_ms}
        engine.profile = True
        run(5)
        engine.profile = False
        ms = engine.kernel_times()
        print(json.dumps({"shape": tag, **{k: round(v, 4) for k, v in ms.items() if v}}), flush=True)

    for key in ("SGBPO_ROLLOUT_R", "SGBPO_ROLLOUT_WARPS", "SGBPO_ROWS_R", "SGBPO_ROWS_WARPS"):
        os.environ.pop(key, None)
    measure("default")
    for shp in a.shapes.split(","):
        r, w = shp.split("x")
        os.environ["SGBPO_ROLLOUT_R"], os.environ["SGBPO_ROLLOUT_WARPS"] = r, w
        measure("rollout " + shp)
    os.environ.pop("SGBPO_ROLLOUT_R"), os.environ.pop("SGBPO_ROLLOUT_WARPS")
    for shp in a.rows.split(","):
        r, w = shp.split("x")
        os.environ["SGBPO_ROWS_R"], os.environ["SGBPO_ROWS_WARPS"] = r, w
        measure("rows " + shp)


if __name__ == "__main__":
    main()
