"""Golden vectors for program P1 on BalanceQuadrotor (wide safe-state generator, 6 x 24): inputs near the boundary of
the RCI set with actions outside the safe action set, and the ORACLE's projection + gradients (oracle/programs.py:
lifted QP, scipy SLSQP + active-set polish; ~4 s per environment, hence a committed fixture instead of a run-time call).

    python tests/golden/make_quadrotor_p1.py            # P1 (boundary projection) -> quadrotor_p1_oracle.npz
    python tests/golden/make_quadrotor_p1.py ray_mask   # orthogonal ray mask (P1 + P4, ray_mask.py:268-341) -> quadrotor_rm_orth_oracle.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE.parent.parent), str(HERE.parent)]
from common import make_case, oracle_step  # noqa: E402

RM = len(sys.argv) > 1 and sys.argv[1] == "ray_mask"
N = 24 if RM else 48
case = make_case("BalanceQuadrotor", "ray_mask", N, seed=7, gate="always", zonotopic=False) if RM else \
    make_case("BalanceQuadrotor", "boundary_projection", N, seed=7, gate="always")
env = case["env"]
d = np.load(HERE.parent.parent / "safegbpo_b200" / "assets" / "assets.npz")
G = torch.from_numpy(d["quadrotor_rci_generators"])
g = torch.Generator().manual_seed(1)
xi = torch.rand(N, 24, generator=g, dtype=torch.float64) * 2 - 1
xi = torch.where(torch.rand(N, 24, generator=g) < 0.5, torch.sign(xi), xi)          # many generators at their bounds
case["s"] = (xi @ G.T) * 0.985
env.state = case["s"].clone()
env.goal = case["s"][:, :2].clone()
case["aux"] = env.goal.clone()
# two thirds of the actions near the corners of the unit box (outside the safe action set / the state-safe region)
corner = torch.sign(case["a"]) * (0.8 + 0.19 * torch.rand(N, 2, generator=g, dtype=torch.float64))
case["a"] = torch.where((torch.arange(N) % 3 != 0).unsqueeze(1), corner, case["a"])
ref = oracle_step(case)

# independent check of the oracle's answers (scipy-HiGHS LP on the UNMERGED encoding): minimal row budget of the
# containment encoding at an action; <= 1 means state-safe.  The oracle's SLSQP stage occasionally accepts a point that
# violates the encoding by a few 1e-3: those environments are dropped from the comparison mask.
from common import containment_budget  # noqa: E402
sg = case["sg"]
env.state = case["s"].clone()
const, b_mat = sg._linear()
Gn, cs = d["quadrotor_rci_generators"], d["quadrotor_rci_center"]
Gw = sg.G_w.numpy()
W = Gw[:, np.abs(Gw).sum(0) > 0]
c_f = np.stack([sg._sets(i, const, b_mat).c_f.detach().numpy() for i in range(N)])
Bm = np.stack([sg._sets(i, const, b_mat).B.detach().numpy() for i in range(N)])


def row_budget(i, act):
    return containment_budget(Gn, W, cs - c_f[i] - Bm[i] @ act)


budget_raw = np.array([row_budget(i, case["a"][i].numpy()) for i in range(N)])
budget_ref = np.array([row_budget(i, ref["a_exec"][i].numpy()) if bool(ref["ok"][i]) else np.inf for i in range(N)])
if not RM:  # (the ray mask's radial map as shipped, quirk Q2, does not keep the action inside the safe set: no budget mask there)
    ref["ok"] = ref["ok"] & torch.from_numpy(budget_ref <= 1 + 1e-6)
out = dict(budget_raw=budget_raw, budget_ref=budget_ref, c_f=c_f, B=Bm, W=W, s=case["s"].numpy(), a=case["a"].numpy(), w=case["w"].numpy(), aux=case["aux"].numpy(), ok=ref["ok"].numpy())
out.update({f"cot_{k}": v.numpy() for k, v in case["cot"].items()})
out.update({f"ref_{k}": ref[k].numpy() for k in ("a_exec", "s_next", "obs", "reward", "gs", "ga")})
np.savez_compressed(HERE / ("quadrotor_rm_orth_oracle.npz" if RM else "quadrotor_p1_oracle.npz"), **out)
moved = np.abs(out["ref_a_exec"] - out["a"]).max(1)
print("oracle ok", int(out["ok"].sum()), "of", N, "; projected", int(((moved > 1e-6) & out["ok"]).sum()))
