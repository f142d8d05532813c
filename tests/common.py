"""Shared helpers for the parity tests: seeded inputs, the oracle run, consts construction."""
import numpy as np
import torch

from oracle.dynamics import OracleEnv, SIMS, ENV_TO_SIM
from oracle.safeguard import OracleSafeguard
from safegbpo_b200 import _lib
from safegbpo_b200.consts import build_consts

ENV_IDS = _lib.ENV_IDS
F64 = torch.float64


def seeded_states(env_name: str, N: int, seed: int, spread: float = 0.97) -> torch.Tensor:
    """States across the feasible box, some pushed to its boundary so projections bind."""
    g = torch.Generator().manual_seed(seed)
    sim = SIMS[ENV_TO_SIM[env_name]]
    c = torch.tensor(sim.state_center, dtype=F64)
    h = torch.tensor(sim.state_half, dtype=F64)
    s = c + (torch.rand(N, sim.n, generator=g, dtype=F64) * 2 - 1) * h * spread
    edge = torch.rand(N, sim.n, generator=g) < 0.15
    sign = torch.where(torch.rand(N, sim.n, generator=g) < 0.5, -1.0, 1.0).to(F64)
    return torch.where(edge, c + sign * h * (0.9 + 0.1 * torch.rand(N, sim.n, generator=g, dtype=F64)), s)


def make_case(env_name: str, kind: str, N: int, seed: int, linear=True, zonotopic=True, passthrough=False,
              gate="always", spread=0.97):
    """Returns dict with oracle env/safeguard, consts, and seeded inputs (float64 CPU tensors)."""
    env = OracleEnv(env_name, N, 240)
    env.reset(seed=seed)
    st = seeded_states(env_name, N, seed, spread)
    env.state = st.clone()
    sg = None
    sg_kind = {"none": _lib.SG_NONE, "boundary_projection": _lib.SG_BOUNDARY_PROJECTION, "ray_mask": _lib.SG_RAY_MASK}[kind]
    g = torch.Generator().manual_seed(seed + 17)
    a = torch.rand(N, env.action_dim, generator=g, dtype=F64) * 2 - 1
    w = env.noise_sample()
    dirs = None
    if kind != "none":
        sg = OracleSafeguard(env, kind, linear_projection=linear, zonotopic_approximation=zonotopic,
                             passthrough=passthrough, gate=gate, strict=False)
        sa = [x.numpy() for x in env.safe_action_set()] if env.action_constrained else None
        ss = [x.numpy() for x in env.safe_state_set()] if env.state_constrained else None
        consts = build_consts(env_name, sg_kind, gate_mode=_lib.GATE[gate], rm_linear=linear, rm_zonotopic=zonotopic,
                              rm_passthrough=passthrough, safe_action=sa, safe_state=ss, G_w=sg.G_w.numpy(),
                              B0=sg.B0.numpy())
        if kind == "ray_mask" and zonotopic and env.state_constrained:
            dirs = sg.draw_directions()
    else:
        consts = build_consts(env_name, sg_kind)
    aux = env.goal.clone() if env.sim.name in ("quadrotor", "seeker") else None
    cot = dict(g_s_next=torch.rand(N, env.state_dim, generator=g, dtype=F64),
               g_a_exec=torch.rand(N, env.action_dim, generator=g, dtype=F64),
               g_obs=torch.rand(N, env.observation().shape[1], generator=g, dtype=F64),
               g_reward=torch.rand(N, generator=g, dtype=F64))
    return dict(env=env, sg=sg, consts=consts, s=st, a=a, w=w, dirs=dirs, aux=aux, cot=cot, kind=kind)


def oracle_step(case):
    """Runs the oracle on the case; returns values, gradients of the cotangent functional and masks."""
    env, sg = case["env"], case["sg"]
    state = case["s"].clone().requires_grad_(True)
    act = case["a"].clone().requires_grad_(True)
    env.state = state
    env.steps = 0
    env.scheduled_reset = False
    if sg is None:
        obs, rew, _, _ = env.step(act, noise=case["w"])
        a_exec = act
        ok = torch.ones(state.shape[0], dtype=torch.bool)
        proj = infeas = None
    else:
        obs, rew, _, _ = sg.step(act, directions=case["dirs"], noise=case["w"])
        a_exec = sg.safe_action
        ok = ~sg.infeasible & ~a_exec.detach().isnan().any(dim=1)
        proj, infeas = sg.projectable.clone(), sg.infeasible.clone()
    cot = case["cot"]
    func = ((env.state * cot["g_s_next"]).sum(1) + (a_exec * cot["g_a_exec"]).sum(1) + (obs * cot["g_obs"]).sum(1)
            + rew * cot["g_reward"])[ok].sum()
    gs, ga = torch.autograd.grad(func, (state, act), allow_unused=True)
    return dict(s_next=env.state.detach(), a_exec=a_exec.detach(), obs=obs.detach(), reward=rew.detach(),
                gs=gs, ga=ga if ga is not None else torch.zeros_like(act), ok=ok, projectable=proj, infeasible=infeas)


def shared_for(env, step_index: int):
    """sgbpo_step_shared for the household series at the post-step index."""
    if env.sim.name != "household":
        return None
    sh = _lib.StepShared()
    t = step_index
    vals = torch.cat([env.load[t:t + 5], env.pv[t:t + 5], env.t_out[t:t + 5], env.price[t:t + 5]]).tolist()
    for i, v in enumerate(vals):
        sh.series[i] = v
    sh.load_t, sh.pv_t, sh.price_t = float(env.load[t]), float(env.pv[t]), float(env.price[t])
    return sh


def p2_center_unique(case, tol: float = 1e-7):
    """P2 (ray-mask expansion) fixes the lengths uniquely (strictly concave objective) but NOT always the centre:
    when the active rows leave a direction of c free, every point of that face is optimal and solvers differ
    (compare only the unique quantities).  For m = 2 with an invertible safe-state generator: range
    of each centre coordinate over the rows with the oracle's optimal lengths, by LP.  Returns a bool mask."""
    from scipy.optimize import linprog
    import oracle.programs as P
    sg, k, dirs = case["sg"], case["consts"], case["dirs"]
    env = case["env"]
    state = env.state
    env.state = case["s"].clone()
    const, b_mat = sg._linear()
    env.state = state
    n = env.state_dim
    Ginv = np.array([k.Ginv[i * n + j] for i in range(n) for j in range(n)]).reshape(n, n)
    cs_hat, rb = np.array(k.cs_hat[:n]), np.array(k.rb[:n])
    B0hat = np.array([k.B0hat[i * 2 + q] for i in range(n) for q in range(2)]).reshape(n, 2)
    mask = np.zeros(case["s"].shape[0], dtype=bool)
    for i in range(case["s"].shape[0]):
        sets = sg._sets(i, const, b_mat)
        try:
            _, ln = P.p2_expansion(dirs[i], sets, sg.B0)
        except P.Infeasible:
            continue
        D, ln = dirs[i].numpy(), ln.detach().numpy()
        cf, B = sets.c_f.detach().numpy(), sets.B.detach().numpy()
        A, g = [], []
        for q in range(2):
            for sgn in (1.0, -1.0):
                a = np.zeros(2); a[q] = sgn
                A.append(a); g.append(1.0 - np.abs(D[q]) @ ln)
        e, b, Wm = cs_hat - Ginv @ cf, Ginv @ B, np.abs(B0hat @ D)
        for r in range(n):
            A.append(-b[r]); g.append(rb[r] - e[r] - Wm[r] @ ln)
            A.append(b[r]); g.append(rb[r] + e[r] - Wm[r] @ ln)
        A, g = np.array(A), np.array(g) + 1e-9
        width = 0.0
        for q in range(2):
            cost = np.zeros(2); cost[q] = 1.0
            lo = linprog(cost, A_ub=A, b_ub=g, bounds=[(None, None)] * 2, method="highs")
            hi = linprog(-cost, A_ub=A, b_ub=g, bounds=[(None, None)] * 2, method="highs")
            if lo.status != 0 or hi.status != 0:
                width = np.inf
                break
            width = max(width, -hi.fun - lo.fun)
        mask[i] = width < tol
    return mask


def containment_budget(G, W, d):
    """Independent feasibility measure for the zonotope-in-zonotope encoding with a wide generator (n x p, p > n):
    the smallest t with  G beta = d,  G Gamma_j = W[:, j],  |beta|_i + sum_j |Gamma_ij| <= t  (scipy HiGHS LP on the
    UNMERGED encoding the reference states, src/safeguards/boundary_projection.py:64-84).  t <= 1 <=> the reachable
    set  d-shifted + W  fits inside the safe set."""
    from scipy.optimize import linprog
    n, p = G.shape
    q = W.shape[1]
    nb, ng = p, q * p
    nvar = 2 * (nb + ng) + 1
    c = np.zeros(nvar)
    c[-1] = 1
    Aeq, beq = np.zeros((n * (1 + q), nvar)), np.zeros(n * (1 + q))
    Aeq[:n, :nb], beq[:n] = G, d
    for j in range(q):
        Aeq[n + n * j:2 * n + n * j, nb + j * p: nb + (j + 1) * p], beq[n + n * j:2 * n + n * j] = G, W[:, j]
    rows = []
    for kx in range(nb + ng):
        for sgn in (1, -1):
            r = np.zeros(nvar)
            r[kx], r[nb + ng + kx] = sgn, -1
            rows.append(r)
    for ii in range(p):
        r = np.zeros(nvar)
        r[nb + ng + ii] = 1
        for j in range(q):
            r[nb + ng + nb + j * p + ii] = 1
        r[-1] = -1
        rows.append(r)
    res = linprog(c, A_ub=np.array(rows), b_ub=np.zeros(len(rows)), A_eq=Aeq, b_eq=beq, bounds=[(None, None)] * nvar,
                  method="highs")
    return res.fun if res.status == 0 else np.inf


def p1_feasibility_budget(G, W, c_s, c_f, B, act_rows=None):
    """Independent feasibility verdict of program P1 for ONE environment (scipy HiGHS LP on the UNMERGED containment
    encoding, src/safeguards/interfaces/safeguard.py:98-197): the smallest t for which some action a in [-1, 1]^m
    (and inside the safe action polytope `act_rows`: rows (n_x, n_y, h) with n . a <= h) has
        G beta = c_s - c_f - B a,   G Gamma_j = W[:, j],   |beta|_i + sum_j |Gamma_ij| <= t.
    t <= 1  <=>  the program is feasible.  G: safe-state generator (n x p), W: noise generator (n x q)."""
    from scipy.optimize import linprog
    n, p = G.shape
    q, m = W.shape[1], B.shape[1]
    nb, ng = p, q * p
    nx = nb + ng
    nvar = m + 2 * nx + 1                      # a | beta, Gamma | their absolute values | t
    c = np.zeros(nvar)
    c[-1] = 1
    Aeq, beq = np.zeros((n * (1 + q), nvar)), np.zeros(n * (1 + q))
    Aeq[:n, :m], Aeq[:n, m:m + nb], beq[:n] = B, G, c_s - c_f
    for j in range(q):
        Aeq[n + n * j:2 * n + n * j, m + nb + j * p: m + nb + (j + 1) * p], beq[n + n * j:2 * n + n * j] = G, W[:, j]
    rows, rhs = [], []
    for kx in range(nx):
        for sgn in (1, -1):
            r = np.zeros(nvar)
            r[m + kx], r[m + nx + kx] = sgn, -1
            rows.append(r); rhs.append(0.0)
    for ii in range(p):
        r = np.zeros(nvar)
        r[m + nx + ii] = 1
        for j in range(q):
            r[m + nx + nb + j * p + ii] = 1
        r[-1] = -1
        rows.append(r); rhs.append(0.0)
    for row in (act_rows if act_rows is not None else []):
        r = np.zeros(nvar)
        r[:m] = row[:m]
        rows.append(r); rhs.append(row[m])
    bounds = [(-1.0, 1.0)] * m + [(None, None)] * (nvar - m)
    res = linprog(c, A_ub=np.array(rows), b_ub=np.array(rhs), A_eq=Aeq, b_eq=beq, bounds=bounds, method="highs")
    return res.fun if res.status == 0 else np.inf
