"""ctypes binding of the C ABI in include/sgbpo.h (libsgbpo.so, hand-written sm_100a CUDA).

There is NO CPU fallback: if the library is missing, cannot be loaded, or a call fails, the product
raises.  ``build()`` compiles the library in-tree with nvcc (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import torch

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_PATH = PKG / "libsgbpo.so"
MAXN, MAXM, MAXHP = 6, 2, 48

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "--shared", "-Xcompiler", "-fPIC"]


class SgbpoError(RuntimeError):
    pass


class Consts(C.Structure):
    """Mirror of ``struct sgbpo_consts`` (include/sgbpo.h)."""
    _fields_ = [
        ("sg_kind", C.c_int32), ("gate_mode", C.c_int32),
        ("rm_linear", C.c_int32), ("rm_zonotopic", C.c_int32), ("rm_passthrough", C.c_int32),
        ("action_constrained", C.c_int32), ("state_constrained", C.c_int32), ("state_model", C.c_int32),
        ("disable_clamp", C.c_int32), ("reserved0", C.c_int32),
        ("box_min", C.c_double * MAXN), ("box_max", C.c_double * MAXN), ("noise_center", C.c_double * MAXN),
        ("a_lo", C.c_double * MAXM), ("a_hi", C.c_double * MAXM),
        ("n_act_hp", C.c_int32), ("act_hp", (C.c_double * 3) * MAXHP), ("c_a", C.c_double * MAXM),
        ("Ginv", C.c_double * (MAXN * MAXN)), ("cs_hat", C.c_double * MAXN), ("rb", C.c_double * MAXN),
        ("B0hat_abs", C.c_double * (MAXN * MAXM)),
        ("n_k_hp", C.c_int32), ("n_q_hp", C.c_int32),
        ("k_hp", (C.c_double * 2) * MAXHP), ("q_hp", (C.c_double * 3) * MAXHP), ("c_s", C.c_double * MAXN),
    ]


class StepShared(C.Structure):
    """Mirror of ``struct sgbpo_step_shared``."""
    _fields_ = [("series", C.c_double * 20), ("load_t", C.c_double), ("pv_t", C.c_double), ("price_t", C.c_double)]


ENV_IDS = {"BalancePendulum": 0, "BalanceCartPole": 1, "SwingUpCartPole": 2, "BalanceQuadrotor": 3,
           "ManageHousehold": 4, "NavigateSeeker": 5}
SG_NONE, SG_BOUNDARY_PROJECTION, SG_RAY_MASK = 0, 1, 2
GATE = {"reference": 0, "intended": 1, "always": 2}
FL_PROJECTABLE, FL_INFEASIBLE, FL_INTERVENED, FL_NONFINITE = 1, 2, 4, 8
FL_GUARD_USED, FL_ACTION_IN_SET, FL_NEXT_IN_SET, FL_CLAMPED = 16, 32, 64, 128

_VP = C.c_void_p
_SIGNATURES = {
    "sgbpo_version": (C.c_char_p, []),
    "sgbpo_last_error": (C.c_char_p, []),
    "sgbpo_env_dims": (C.c_int, [C.c_int] + [C.POINTER(C.c_int)] * 4),
    "sgbpo_step_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(Consts), C.POINTER(StepShared)]
                       + [_VP] * 6 + [_VP] * 4 + [_VP, _VP]),
    "sgbpo_step_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(Consts), C.POINTER(StepShared)]
                       + [_VP] * 6 + [_VP] * 4 + [_VP, _VP, _VP]),
    "sgbpo_linearise": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(Consts)] + [_VP] * 5 + [_VP]),
}
EXPORTED = tuple(_SIGNATURES)
_lib = None


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def build(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libsgbpo.so (in-tree, so it travels with the repo)."""
    srcs = sources()
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [PKG.parent / "include" / "sgbpo.h"])
    if LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= newest:
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SgbpoError(f"nvcc failed:\n{res.stderr[-4000:]}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libsgbpo.so; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SgbpoError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                             "safegbpo_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        raise SgbpoError(f"{what} failed ({code}): {lib().sgbpo_last_error().decode()}")


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return 0
    if t == torch.float64:
        return 1
    raise SgbpoError(f"unsupported dtype {t}: kernels are instantiated for float32 and float64")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SgbpoError("safegbpo_b200 kernels need CUDA tensors (no CPU path)")
    if not t.is_contiguous():
        raise SgbpoError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
