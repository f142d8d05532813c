"""Tuning aid: wall-clock of the SHAC critic path (TD-lambda targets, critic fit, Polyak) next to the rollout."""
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
    torch.set_default_dtype(torch.float32)
    torch.set_default_device(torch.device("cuda", 0))
    torch.manual_seed(0)
    env = BoundaryProjectionSafeguard(balance_cartpole.BalanceCartPoleEnv(num_envs=4096, num_steps=240), strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": bench.POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": bench.VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=64, fused="on", rng_order="batched")
    marks = {}

    def stamp(name, t0):
        torch.cuda.synchronize()
        marks.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    for ep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        algo.buffer.reset(); algo.policy_optim.zero_grad()
        algo.collect_trajectories(); algo.update_policy(); stamp("rollout+policy", t0)
        t1 = time.perf_counter()
        with torch.no_grad():
            est, obs = algo.calculate_estimated_values()
        stamp("td_lambda_targets", t1)
        t2 = time.perf_counter()
        algo.update_value_function(); stamp("update_value_function (incl. targets)", t2)
        t3 = time.perf_counter()
        algo.update_target_value_function(); stamp("polyak", t3)
    print(json.dumps({k: round(sum(v[3:]) / len(v[3:]), 3) for k, v in marks.items()}))


if __name__ == "__main__":
    main()
