"""torch.autograd bridges over the C ABI (include/sgbpo.h).  PyTorch is plumbing here: it owns the
device memory and the stream; the arithmetic is in libsgbpo.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import Consts, StepShared, check, dtype_code, ptr, stream_ptr

ENV_DIMS = {}


def env_dims(env_id: int):
    """(n, m, o, naux) from the library's own shape table."""
    if env_id not in ENV_DIMS:
        vals = [C.c_int() for _ in range(4)]
        check(_lib.lib().sgbpo_env_dims(env_id, *[C.byref(v) for v in vals]), "sgbpo_env_dims")
        ENV_DIMS[env_id] = tuple(v.value for v in vals)
    return ENV_DIMS[env_id]


def _c(t: Optional[Tensor], dtype=None) -> Optional[Tensor]:
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class SafeguardedStepFn(torch.autograd.Function):
    """One fused safeguarded environment step.

    Replaces, per call: Safeguard.actions (safeguards/interfaces/safeguard.py:61-80: linearise, gate,
    CvxpyLayer solve, where) + Simulator.execute_action / observation / reward
    (envs/simulators/interfaces/simulator.py:82-108,174-182).
    Inputs: state (N,n), action (N,m) [differentiable]; noise (N,n), aux (N,naux), dirs (N,m,2m),
    reset_state (N,n) [constants].  Outputs: next_state, exec_action, obs, reward, flags(int32).
    """

    @staticmethod
    def forward(ctx, state, action, noise, aux, dirs, reset_state, env_id, consts, shared):
        n, m, o, naux = env_dims(env_id)
        dt = state.dtype
        state, action, noise = _c(state), _c(action, dt), _c(noise, dt)
        aux, dirs, reset_state = _c(aux, dt), _c(dirs, dt), _c(reset_state, dt)
        N = state.shape[0]
        if state.shape != (N, n) or action.shape != (N, m) or noise.shape != (N, n):
            raise _lib.SgbpoError(f"shape mismatch: state {tuple(state.shape)} action {tuple(action.shape)} "
                                  f"noise {tuple(noise.shape)} for env dims n={n} m={m}")
        if naux and (aux is None or aux.shape != (N, naux)):
            raise _lib.SgbpoError("this environment needs aux (N, naux)")
        opts = dict(device=state.device, dtype=dt)
        s_next = torch.empty(N, n, **opts)
        a_exec = torch.empty(N, m, **opts)
        obs = torch.empty(N, o, **opts)
        reward = torch.empty(N, **opts)
        flags = torch.empty(N, device=state.device, dtype=torch.int32)
        check(_lib.lib().sgbpo_step_fwd(env_id, dtype_code(dt), N, C.byref(consts),
                                        C.byref(shared) if shared is not None else None,
                                        ptr(state), ptr(action), ptr(noise), ptr(aux), ptr(dirs), ptr(reset_state),
                                        ptr(s_next), ptr(a_exec), ptr(obs), ptr(reward), ptr(flags), stream_ptr()),
              "sgbpo_step_fwd")
        ctx.save_for_backward(state, action, noise, aux, dirs, reset_state)
        ctx.meta = (env_id, consts, shared)
        ctx.mark_non_differentiable(flags)
        return s_next, a_exec, obs, reward, flags

    @staticmethod
    def backward(ctx, g_s_next, g_a_exec, g_obs, g_reward, _g_flags):
        state, action, noise, aux, dirs, reset_state = ctx.saved_tensors
        env_id, consts, shared = ctx.meta
        dt = state.dtype
        N = state.shape[0]
        gs, ga = torch.empty_like(state), torch.empty_like(action)
        g = [None if x is None else _c(x, dt) for x in (g_s_next, g_a_exec, g_obs, g_reward)]
        check(_lib.lib().sgbpo_step_bwd(env_id, dtype_code(dt), N, C.byref(consts),
                                        C.byref(shared) if shared is not None else None,
                                        ptr(state), ptr(action), ptr(noise), ptr(aux), ptr(dirs), ptr(reset_state),
                                        ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(gs), ptr(ga), stream_ptr()),
              "sgbpo_step_bwd")
        return gs, ga, None, None, None, None, None, None, None


def safeguarded_step(state: Tensor, action: Tensor, noise: Tensor, env_id: int, consts: Consts,
                     aux: Optional[Tensor] = None, dirs: Optional[Tensor] = None,
                     reset_state: Optional[Tensor] = None, shared: Optional[StepShared] = None):
    return SafeguardedStepFn.apply(state, action, noise, aux, dirs, reset_state, env_id, consts, shared)


def linearise(state: Tensor, env_id: int, consts: Consts):
    """Simulator.linear_dynamics (simulator.py:206-230): (c, A, B, E).  Forward only: gradients w.r.t.
    the state flow through the fused step op, which differentiates the linearisation internally."""
    n, m, o, naux = env_dims(env_id)
    state = _c(state.detach())
    N = state.shape[0]
    opts = dict(device=state.device, dtype=state.dtype)
    c, A = torch.empty(N, n, **opts), torch.empty(N, n, n, **opts)
    B, E = torch.empty(N, n, m, **opts), torch.empty(N, n, n, **opts)
    check(_lib.lib().sgbpo_linearise(env_id, dtype_code(state.dtype), N, C.byref(consts), ptr(state), ptr(c), ptr(A),
                                     ptr(B), ptr(E), stream_ptr()), "sgbpo_linearise")
    return c, A, B, E
