"""ctypes access to tests/hostmath/libhostmath.so (the device math compiled for the host; test-only)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from safegbpo_b200._lib import Consts, StepShared

HERE = Path(__file__).resolve().parent / "hostmath"
DIMS = {0: (2, 1, 3, 0), 1: (4, 1, 5, 0), 2: (4, 1, 5, 0), 3: (6, 2, 8, 2), 4: (3, 2, 23, 0), 5: (2, 2, 4, 11), 6: (6, 2, 8, 13)}
_lib = None


def lib():
    global _lib
    if _lib is None:
        so, src = HERE / "libhostmath.so", HERE / "hostmath.cpp"
        deps = [src] + list((HERE.parent.parent / "safegbpo_b200" / "csrc").glob("*.cuh"))
        if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in deps):
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(so), str(src)],
                           check=True)
        _lib = C.CDLL(str(so))
        _lib.hm_step.restype = C.c_int
        _lib.hm_linearise.restype = C.c_int
    return _lib


def _p(x):
    return None if x is None else x.ctypes.data_as(C.c_void_p)


def step(env_id, consts: Consts, s, a, w, aux=None, dirs=None, reset_state=None, shared=None, cot=None, dtype=np.float64):
    """Forward (+ backward when ``cot`` = dict(g_s_next,g_a_exec,g_obs,g_reward)) on the host.  Returns dict."""
    n, m, o, naux = DIMS[env_id]
    cast = lambda x: None if x is None else np.ascontiguousarray(x, dtype=dtype)
    s, a, w, aux, dirs, reset_state = map(cast, (s, a, w, aux, dirs, reset_state))
    N = s.shape[0]
    out = dict(s_next=np.zeros((N, n), dtype), a_exec=np.zeros((N, m), dtype), obs=np.zeros((N, o), dtype),
               reward=np.zeros(N, dtype), flags=np.zeros(N, np.int32))
    g = {k: cast(v) for k, v in (cot or {}).items()}
    gs = np.zeros((N, n), dtype) if cot is not None else None
    ga = np.zeros((N, m), dtype) if cot is not None else None
    rc = lib().hm_step(C.c_int(env_id), C.c_int(1 if dtype == np.float64 else 0), C.c_int64(N), C.byref(consts),
                       C.byref(shared) if shared is not None else None, _p(s), _p(a), _p(w), _p(aux), _p(dirs),
                       _p(reset_state), _p(out["s_next"]), _p(out["a_exec"]), _p(out["obs"]), _p(out["reward"]),
                       _p(out["flags"]), _p(g.get("g_s_next")), _p(g.get("g_a_exec")), _p(g.get("g_obs")),
                       _p(g.get("g_reward")), _p(gs), _p(ga))
    assert rc == 0
    out["gs"], out["ga"] = gs, ga
    return out


def linearise(env_id, consts: Consts, s, dtype=np.float64):
    n, m, o, naux = DIMS[env_id]
    s = np.ascontiguousarray(s, dtype=dtype)
    N = s.shape[0]
    c, A, B, E = np.zeros((N, n), dtype), np.zeros((N, n, n), dtype), np.zeros((N, n, m), dtype), np.zeros((N, n, n), dtype)
    rc = lib().hm_linearise(C.c_int(env_id), C.c_int(1 if dtype == np.float64 else 0), C.c_int64(N), C.byref(consts),
                            _p(s), _p(c), _p(A), _p(B), _p(E))
    assert rc == 0
    return c, A, B, E


def navquad_safe_state_set(consts: Consts, s, aux, dtype=np.float64):
    """NavigateQuadrotor's per-step safe state set (P6-P8) on the host: (center (N,6), generator (N,6,6), status (N))."""
    s, aux = np.ascontiguousarray(s, dtype=dtype), np.ascontiguousarray(aux, dtype=dtype)
    N = s.shape[0]
    center, gen, status = np.zeros((N, 6), dtype), np.zeros((N, 6, 6), dtype), np.zeros(N, np.int32)
    lib().hm_navquad_set.restype = C.c_int
    rc = lib().hm_navquad_set(C.c_int(1 if dtype == np.float64 else 0), C.c_int64(N), C.byref(consts), _p(s), _p(aux), _p(center),
                              _p(gen), _p(status))
    assert rc == 0
    return center, gen, status


def wrap_angle(x, dtype=np.float64):
    """csrc/sgbpo_dual.cuh::wrap_angle on the host: (wrapped value, derivative)."""
    x = np.ascontiguousarray(x, dtype=dtype)
    y, dy = np.empty_like(x), np.empty_like(x)
    lib().hm_wrap_angle.restype = C.c_int
    lib().hm_wrap_angle(int(dtype == np.float64), C.c_int64(x.size), _p(x), _p(y), _p(dy))
    return y, dy
