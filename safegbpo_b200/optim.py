"""Fused tail of ``SHAC.update_policy`` (reference: src/learning_algorithms/shac.py:148-149):
``clip_grad_norm_(policy.parameters(), 1.0)`` followed by ``policy_optim.step()`` as ONE kernel launch
(csrc/sgbpo_optim.cu) working directly on torch.optim.Adam's own state tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Optim(C.Structure):
    """Mirror of ``struct sgbpo_optim`` (include/sgbpo.h)."""
    MAX = 24
    _fields_ = [("n_tensors", C.c_int32), ("reserved", C.c_int32),
                ("param", C.c_void_p * 24), ("grad", C.c_void_p * 24), ("exp_avg", C.c_void_p * 24),
                ("exp_avg_sq", C.c_void_p * 24), ("step", C.c_void_p * 24), ("numel", C.c_int64 * 24),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("max_norm", C.c_double)]


def supported(optim: torch.optim.Optimizer) -> bool:
    """Plain Adam on CUDA float32/float64 parameters with device-side step counters (capturable=True)."""
    if type(optim) is not torch.optim.Adam or len(optim.param_groups) != 1:
        return False
    g = optim.param_groups[0]
    if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False) or not g.get("capturable", False):
        return False
    if isinstance(g["lr"], torch.Tensor) or g.get("decoupled_weight_decay", False):
        return False
    ps = g["params"]
    if not ps or len(ps) > Optim.MAX:
        return False
    dt = ps[0].dtype
    return all(p.is_cuda and p.dtype == dt and p.is_contiguous() for p in ps) and dt in (torch.float32, torch.float64)


def _ensure_state(optim: torch.optim.Adam):
    """What Adam._init_group creates lazily on the first step()."""
    for p in optim.param_groups[0]["params"]:
        st = optim.state[p]
        if len(st) == 0:
            step_dtype = torch.float64 if torch.get_default_dtype() == torch.float64 else torch.float32
            st["step"] = torch.zeros((), dtype=step_dtype, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)


def clip_adam_step(optim: torch.optim.Adam, max_norm: float) -> torch.Tensor:
    """clip_grad_norm_ + Adam.step in one launch; returns the total gradient norm (0-dim device tensor).
    Parameters without a gradient are skipped, as torch does.  The argument block is cached while the
    gradient, parameter and optimizer-state tensors stay the same storage (the fused engine re-attaches its static
    gradient buffers every episode)."""
    g = optim.param_groups[0]
    ps = [p for p in g["params"] if p.grad is not None]
    _ensure_state(optim)
    # every pointer the argument block bakes in: gradients, parameters AND the optimizer's state tensors (load_state_dict(),
    # .to() or a checkpoint resume replace them without touching the gradients)
    key = tuple((p.grad.data_ptr(), p.data_ptr(), optim.state[p]["exp_avg"].data_ptr(), optim.state[p]["exp_avg_sq"].data_ptr(),
                 optim.state[p]["step"].data_ptr()) for p in ps)
    cache = getattr(optim, "_sgbpo_cache", None)
    if cache is None or cache[0] != key:
        o = Optim()
        o.n_tensors = len(ps)
        step_dtype = None
        for i, p in enumerate(ps):
            st = optim.state[p]
            grad = p.grad
            if not grad.is_contiguous() or grad.dtype != p.dtype:
                raise _lib.SgbpoError("sgbpo_clip_adam: gradients must be contiguous and of the parameter dtype")
            o.param[i], o.grad[i] = p.data_ptr(), grad.data_ptr()
            o.exp_avg[i], o.exp_avg_sq[i], o.step[i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr()
            o.numel[i] = p.numel()
            if not st["step"].is_cuda:
                raise _lib.SgbpoError("sgbpo_clip_adam needs device-side step counters (Adam(capturable=True))")
            sd = st["step"].dtype
            if step_dtype is not None and sd != step_dtype:
                raise _lib.SgbpoError("sgbpo_clip_adam: mixed step-counter dtypes")
            step_dtype = sd
        total = torch.empty((), dtype=ps[0].dtype, device=ps[0].device)
        keep = [optim.state[p] for p in ps]        # the state tensors whose pointers the block holds
        cache = (key, o, total, _lib.dtype_code(ps[0].dtype), _lib.dtype_code(step_dtype), keep)
        optim._sgbpo_cache = cache
    _, o, total, dt, sdt, _ = cache
    beta1, beta2 = g["betas"]
    o.lr, o.beta1, o.beta2, o.eps, o.max_norm = float(g["lr"]), float(beta1), float(beta2), float(g["eps"]), float(max_norm)
    _lib.check(_lib.lib().sgbpo_clip_adam(dt, sdt, C.byref(o), C.c_void_p(total.data_ptr()), _lib.stream_ptr()), "sgbpo_clip_adam")
    return total
