"""Time the per-environment lifted P1 solve (csrc/sgbpo_lifted.cuh) on BalanceQuadrotor: every environment solves
(gate="always"), states spread over the RCI set, actions near the corners of the action box.

    python tools/lifted_time.py [--num-envs 4096] [--dtype f32|f64]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from safegbpo_b200 import _lib, ops  # noqa: E402
from safegbpo_b200.envs.balance_quadrotor import BalanceQuadrotorEnv  # noqa: E402
from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--scale", type=float, default=0.97)
    args = ap.parse_args()
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    dev = torch.device("cuda:0")
    torch.set_default_dtype(dtype)
    torch.set_default_device(dev)
    N = args.num_envs
    env = BalanceQuadrotorEnv(num_envs=N, num_steps=240)
    sg = BoundaryProjectionSafeguard(env, gate="always", strict=False)
    sg.reset(seed=0)
    d = np.load(Path(_lib.__file__).parent / "assets" / "assets.npz")
    G = torch.from_numpy(d["quadrotor_rci_generators"]).to(dev, torch.float64)
    g = torch.Generator(device="cpu").manual_seed(1)
    xi = (torch.rand(N, 24, generator=g, dtype=torch.float64, device="cpu") * 2 - 1).to(dev)
    s = ((xi @ G.T) * args.scale).to(dtype)
    a = (torch.sign(torch.rand(N, 2, generator=g, device="cpu") - 0.5) * (0.8 + 0.19 * torch.rand(N, 2, generator=g, device="cpu"))).to(dev, dtype)
    w = torch.zeros(N, 6)
    aux = s[:, :2].clone()
    k = sg.consts()

    def fwd():
        return ops.safeguarded_step(s, a, w, env.env_id, k, aux=aux)

    def both():
        s_, a_ = s.clone().requires_grad_(True), a.clone().requires_grad_(True)
        out = ops.safeguarded_step(s_, a_, w, env.env_id, k, aux=aux)
        (out[0].sum() + out[1].sum()).backward()
        return out

    for name, fn in (("forward", fwd), ("forward+backward", both)):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        fl = out[4]
        moved = (out[1].detach() - a).abs().amax(1) > 1e-6
        print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per step of {N} envs ({args.dtype}); guarded "
              f"{int(((fl & _lib.FL_GUARD_USED) != 0).sum())}, moved {int(moved.sum())}, infeasible "
              f"{int(((fl & _lib.FL_INFEASIBLE) != 0).sum())}")


if __name__ == "__main__":
    main()
