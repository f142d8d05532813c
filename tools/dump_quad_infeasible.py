"""Debug aid: run the system-matrix BalanceQuadrotor + BP SHAC episodes and dump the (state, action) pairs whose program
was reported infeasible -> gpurun_out/quad_infeasible.npz (analysed on the CPU against an independent LP)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from safegbpo_b200 import _lib  # noqa: E402
from safegbpo_b200.envs.balance_quadrotor import BalanceQuadrotorEnv  # noqa: E402
from safegbpo_b200.learning_algorithms.shac import SHAC  # noqa: E402
from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard  # noqa: E402

dtype = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
torch.set_default_dtype(dtype)
torch.set_default_device("cuda:0")
torch.manual_seed(1)
if len(sys.argv) > 2 and sys.argv[2] == "rm_orthogonal_always":
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    env = RayMaskSafeguard(BalanceQuadrotorEnv(num_envs=16, num_steps=240), regularisation_coefficient=0.1, linear_projection=True,
                           zonotopic_approximation=False, passthrough=False, gate="always")
else:
    env = BoundaryProjectionSafeguard(BalanceQuadrotorEnv(num_envs=16, num_steps=240))
algo = SHAC(env=env, policy_kwargs={"net_arch": [64, 64]}, policy_optim_kwargs={"lr": 1e-3},
            vf_kwargs={"net_arch": [64, 64]}, vf_optim_kwargs={"lr": 1e-3}, regularisation_coefficient=0.1,
            len_trajectories=8, vf_num_fits=2, vf_fit_num_batches=2)
rows = []
for eps in range(12):
    try:
        algo._learn_episode(eps)
    except ValueError as exc:
        print("episode", eps, "raised:", str(exc)[:60])
        eng = algo._fused_engine
        g = eng._g if (eng.last or {}).get("graph") else dict(eng.last, H=eng.last["flags"].shape[0], actions=algo.buffer.actions.tensor)
        H = g["H"]
        fl = g["flags"][:H]
        bad = (fl & _lib.FL_INFEASIBLE) != 0
        print("infeasible entries:", bad.nonzero().tolist(), [hex(int(x)) for x in fl[bad].tolist()])
        nan = g["exec_actions"][:H].isnan().any(-1)
        print("nan actions:", nan.nonzero().tolist())
        sel = bad | nan
        first = int(sel.any(1).nonzero()[0])
        print("first bad step", first, "flags there", [hex(int(x)) for x in fl[first].tolist()])
        sel[:] = False
        sel[first] = bad[first] | nan[first]
        np.savez(Path(__file__).resolve().parent.parent / "gpurun_out" / "quad_infeasible.npz",
                 s=g["states"][:H][sel].double().cpu().numpy(), a=g["actions"][:H][sel].double().cpu().numpy(),
                 flags=fl[sel].cpu().numpy(), states=g["states"].double().cpu().numpy(), flags_all=fl.cpu().numpy())
        break
else:
    print("no infeasible program in 12 episodes")
