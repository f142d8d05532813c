"""Seeded NavigateSeeker cases shared by the CPU (host-math) and GPU tests."""
import torch

from oracle.dynamics import OracleEnv
from oracle.safeguard import OracleSafeguard
from safegbpo_b200 import _lib
from safegbpo_b200.consts import build_consts

F64 = torch.float64


def aux_of(env, N):
    aux = torch.zeros(N, 11, dtype=F64)
    aux[:, :2] = env.goal
    for j, (c, r) in enumerate(env.obstacles):
        aux[:, 2 + 3 * j:4 + 3 * j] = c
        aux[:, 4 + 3 * j] = r
    return aux


def safeguard_case(kind, linear, N=14, seed=0, nobs=2):
    g = torch.Generator().manual_seed(seed)
    env = OracleEnv("NavigateSeeker", N, 400)
    st = (torch.rand(N, 2, generator=g, dtype=F64) * 2 - 1) * 7.5
    env.goal = (torch.rand(N, 2, generator=g, dtype=F64) * 2 - 1) * 7.5
    obstacles = []
    for _ in range(nobs):
        ang = torch.rand(N, generator=g, dtype=F64) * 6.28
        dist = 2.0 + torch.rand(N, generator=g, dtype=F64) * 0.6
        rad = 1.5 + torch.rand(N, generator=g, dtype=F64) * 0.45
        obstacles.append((st + torch.stack([ang.cos(), ang.sin()], 1) * dist.unsqueeze(1), rad))
    env.obstacles = obstacles
    env.state, env.scheduled_reset = st.clone(), False
    sg = OracleSafeguard(env, kind, linear_projection=linear, gate="always", strict=False)
    sgk = _lib.SG_BOUNDARY_PROJECTION if kind == "boundary_projection" else _lib.SG_RAY_MASK
    consts = build_consts("NavigateSeeker", sgk, gate_mode=2, rm_linear=linear, dynamic_action_set=True, n_obstacles=nobs,
                          G_w=sg.G_w.numpy(), B0=sg.B0.numpy())
    a = torch.rand(N, 2, generator=g, dtype=F64) * 2 - 1
    w = env.noise_sample()
    state, act = st.clone().requires_grad_(True), a.clone().requires_grad_(True)
    env.state = state
    obs, rew, _, _ = sg.step(act, noise=w)
    cot = dict(g_s_next=torch.rand(N, 2, generator=g, dtype=F64), g_a_exec=torch.rand(N, 2, generator=g, dtype=F64),
               g_reward=torch.rand(N, generator=g, dtype=F64))
    ok = ~sg.infeasible & ~sg.safe_action.detach().isnan().any(1)
    func = ((env.state * cot["g_s_next"]).sum(1) + (sg.safe_action * cot["g_a_exec"]).sum(1) + rew * cot["g_reward"])[ok].sum()
    gs, ga = torch.autograd.grad(func, (state, act))
    ref = dict(a_exec=sg.safe_action.detach(), s_next=env.state.detach(), reward=rew.detach(), gs=gs, ga=ga, ok=ok,
               generator=sg._action_set[1])
    return dict(consts=consts, s=st, a=a, w=w, aux=aux_of(env, N), cot=cot, ref=ref)


def collision_case(N=16, seed=1):
    g = torch.Generator().manual_seed(seed)
    env = OracleEnv("NavigateSeeker", N, 400)
    st = (torch.rand(N, 2, generator=g, dtype=F64) * 2 - 1) * 6
    env.goal = (torch.rand(N, 2, generator=g, dtype=F64) * 2 - 1) * 7
    ang = torch.rand(N, generator=g, dtype=F64) * 6.28
    rad = 2 + torch.rand(N, generator=g, dtype=F64)
    dirn = torch.stack([ang.cos(), ang.sin()], 1)
    c0 = st + dirn * (rad + 0.03).unsqueeze(1)
    c1 = st - dirn * (rad + 3.0).unsqueeze(1)
    env.obstacles = [(c0, rad), (c1, rad.clone())]
    a = dirn * 0.9 + (torch.rand(N, 2, generator=g, dtype=F64) - 0.5) * 0.4
    a[N // 2:] = -a[N // 2:]
    w = env.noise_sample() * 0.2
    state, act = st.clone().requires_grad_(True), a.clone().requires_grad_(True)
    env.state, env.scheduled_reset = state, False
    obs, rew, _, _ = env.step(act, noise=w)
    cot = dict(g_s_next=torch.rand(N, 2, generator=g, dtype=F64), g_reward=torch.rand(N, generator=g, dtype=F64))
    gs, ga = torch.autograd.grad((env.state * cot["g_s_next"]).sum() + (rew * cot["g_reward"]).sum(), (state, act))
    ref = dict(s_next=env.state.detach(), reward=rew.detach(), gs=gs, ga=ga, collided=env.collided.any(1), obs=obs.detach())
    return dict(consts=build_consts("NavigateSeeker", _lib.SG_NONE, n_obstacles=2), s=st, a=a, w=w, aux=aux_of(env, N), cot=cot, ref=ref)
