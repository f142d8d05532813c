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
MAXGEN, MAXNULL = 20, 14

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
        ("disable_clamp", C.c_int32), ("n_obstacles", C.c_int32), ("act_set_mode", C.c_int32), ("infeasible_raw", C.c_int32),
        ("box_min", C.c_double * MAXN), ("box_max", C.c_double * MAXN), ("noise_center", C.c_double * MAXN),
        ("a_lo", C.c_double * MAXM), ("a_hi", C.c_double * MAXM),
        ("n_act_hp", C.c_int32), ("act_hp", (C.c_double * 3) * MAXHP), ("c_a", C.c_double * MAXM),
        ("Ginv", C.c_double * (MAXN * MAXN)), ("cs_hat", C.c_double * MAXN), ("rb", C.c_double * MAXN),
        ("B0hat_abs", C.c_double * (MAXN * MAXM)),
        ("n_k_hp", C.c_int32), ("n_q_hp", C.c_int32),
        ("k_hp", (C.c_double * 2) * MAXHP), ("q_hp", (C.c_double * 3) * MAXHP), ("c_s", C.c_double * MAXN),
        ("B0hat", C.c_double * (MAXN * MAXM)),
        ("n_gen", C.c_int32), ("n_null", C.c_int32), ("n_wcol", C.c_int32), ("reserved2", C.c_int32),
        ("Gpinv", (C.c_double * MAXN) * MAXGEN), ("Nmat", (C.c_double * MAXNULL) * MAXGEN), ("Gam_p", (C.c_double * 2) * MAXGEN),
    ]


class StepShared(C.Structure):
    """Mirror of ``struct sgbpo_step_shared``."""
    _fields_ = [("series", C.c_double * 20), ("load_t", C.c_double), ("pv_t", C.c_double), ("price_t", C.c_double)]


class Mlp(C.Structure):
    """Mirror of ``struct sgbpo_mlp`` (pointers to the torch parameters themselves, no repacking)."""
    _fields_ = [("in_dim", C.c_int32), ("width", C.c_int32), ("n_hidden", C.c_int32), ("out_dim", C.c_int32),
                ("weight", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("ln_weight", C.c_void_p * 4),
                ("ln_bias", C.c_void_p * 4), ("log_std", C.c_void_p)]


class Rollout(C.Structure):
    """Mirror of ``struct sgbpo_rollout``."""
    _fields_ = [("N", C.c_int64), ("H", C.c_int32), ("engine", C.c_int32),
                ("eps", C.c_void_p), ("noise", C.c_void_p), ("dirs", C.c_void_p), ("aux0", C.c_void_p),
                ("reset_states", C.c_void_p), ("reset_idx", C.c_void_p), ("aux_src", C.c_void_p), ("shared", C.c_void_p),
                ("states", C.c_void_p), ("obs", C.c_void_p), ("actions", C.c_void_p), ("exec_actions", C.c_void_p),
                ("rewards", C.c_void_p), ("flags", C.c_void_p),
                ("stash_h1", C.c_void_p), ("stash_e1", C.c_void_p), ("stash_e2", C.c_void_p), ("stash_stats", C.c_void_p),
                ("reset_aux", C.c_void_p), ("jac", C.c_void_p), ("pol_jac", C.c_void_p), ("dmu", C.c_void_p)]


class RolloutGrads(C.Structure):
    """Mirror of ``struct sgbpo_rollout_grads``."""
    _fields_ = [("grw", C.c_void_p), ("gobs_term", C.c_void_p), ("term_of_row", C.c_void_p),
                ("dz2_out", C.c_void_p), ("g_w_in", C.c_void_p), ("g_w_out", C.c_void_p),
                ("g_b_hid", C.c_void_p), ("g_gamma", C.c_void_p), ("g_beta", C.c_void_p), ("g_b_out", C.c_void_p),
                ("g_log_std", C.c_void_p), ("g_w_hid", C.c_void_p)]


class EpisodeInputs(C.Structure):
    """Mirror of ``struct sgbpo_episode_inputs_args``."""
    _fields_ = [("N", C.c_int64), ("H", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("n_resets", C.c_int32),
                ("n_reset_gens", C.c_int32), ("want_dirs", C.c_int32), ("seed", C.c_uint64), ("episode", C.c_void_p),
                ("eps", C.c_void_p), ("noise", C.c_void_p), ("dirs", C.c_void_p), ("reset_states", C.c_void_p),
                ("noise_lo", C.c_double * MAXN), ("noise_span", C.c_double * MAXN),
                ("reset_center", C.c_double * MAXN), ("reset_gen", (C.c_double * 24) * MAXN)]


class ContainsArgs(C.Structure):
    """Mirror of ``struct sgbpo_contains_args``."""
    _fields_ = [("N", C.c_int64), ("n", C.c_int32), ("K", C.c_int32), ("q", C.c_int32), ("center_per_env", C.c_int32),
                ("facets", C.c_void_p), ("outer_center", C.c_void_p), ("points", C.c_void_p), ("inner_gen", C.c_void_p),
                ("value", C.c_void_p), ("verdict", C.c_void_p), ("tol", C.c_double)]


ENV_IDS = {"BalancePendulum": 0, "BalanceCartPole": 1, "SwingUpCartPole": 2, "BalanceQuadrotor": 3,
           "ManageHousehold": 4, "NavigateSeeker": 5, "NavigateQuadrotor": 6}
SG_NONE, SG_BOUNDARY_PROJECTION, SG_RAY_MASK = 0, 1, 2
GATE = {"reference": 0, "intended": 1, "always": 2}
FL_PROJECTABLE, FL_INFEASIBLE, FL_INTERVENED, FL_NONFINITE = 1, 2, 4, 8
FL_GUARD_USED, FL_ACTION_IN_SET, FL_NEXT_IN_SET, FL_CLAMPED, FL_COLLIDED, FL_UNREDUCED = 16, 32, 64, 128, 256, 512
FL_REACHED = 1024

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
    "sgbpo_seeker_action_set": (C.c_int, [C.c_int, C.c_int64, C.POINTER(Consts)] + [_VP] * 5),
    "sgbpo_navquad_safe_state_set": (C.c_int, [C.c_int, C.c_int64, C.POINTER(Consts)] + [_VP] * 6),
    "sgbpo_rollout_fwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(Consts), C.POINTER(Mlp), C.POINTER(Rollout), _VP]),
    "sgbpo_rollout_bwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(Consts), C.POINTER(Mlp), C.POINTER(Rollout),
                                    C.POINTER(RolloutGrads), _VP]),
    "sgbpo_rollout_jac": (C.c_int, [C.c_int, C.c_int, C.POINTER(Consts), C.POINTER(Rollout), _VP]),
    "sgbpo_rollout_jac_stride": (C.c_int, [C.c_int]),
    "sgbpo_rollout_pol_jac_stride": (C.c_int, [C.c_int]),
    "sgbpo_rollout_engine_supported": (C.c_int, [C.c_int, C.c_int, C.POINTER(Mlp)]),
    "sgbpo_mlp_rows": (C.c_int, [C.c_int, C.c_int64, C.POINTER(Mlp), _VP, _VP, _VP]),
    "sgbpo_mlp_rows_vjp": (C.c_int, [C.c_int, C.c_int64, C.POINTER(Mlp), _VP, _VP, _VP, _VP]),
    "sgbpo_clip_adam": (C.c_int, [C.c_int, C.c_int, _VP, _VP, _VP]),
    "sgbpo_bias_elu_ln_fwd": (C.c_int, [C.c_int, C.c_int64, C.c_int] + [_VP] * 8),
    "sgbpo_bias_elu_ln_bwd": (C.c_int, [C.c_int, C.c_int64, C.c_int] + [_VP] * 9),
    "sgbpo_td_lambda": (C.c_int, [C.c_int, C.c_int64, C.c_int, _VP, _VP, _VP, C.c_double, C.c_double, _VP, _VP, _VP]),
    "sgbpo_episode_stats": (C.c_int, [C.c_int, C.c_int64, _VP, C.c_int64, _VP, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "sgbpo_shac_loss": (C.c_int, [C.c_int, C.c_int, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "sgbpo_tc_selftest": (C.c_int, [C.c_int, _VP, _VP, _VP, _VP]),
    "sgbpo_episode_inputs": (C.c_int, [C.c_int, C.POINTER(EpisodeInputs), _VP]),
    "sgbpo_contains": (C.c_int, [C.c_int, C.POINTER(ContainsArgs), _VP]),
}
EXPORTED = tuple(_SIGNATURES)
_lib = None


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def build(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libsgbpo.so (in-tree, so it travels with the repo).
    One object per translation unit, compiled in parallel, then linked."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sources()
    import hashlib
    import re

    inc_re = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    dep_cache: dict[Path, tuple[Path, ...]] = {}

    def deps(path: Path) -> tuple[Path, ...]:
        """Transitive closure of the quoted includes of `path` (resolved relative to the including file)."""
        path = path.resolve()
        if path in dep_cache:
            return dep_cache[path]
        dep_cache[path] = ()
        found: list[Path] = []
        for rel in inc_re.findall(path.read_text()):
            inc = (path.parent / rel).resolve()
            if inc.exists():
                found.append(inc)
                found.extend(deps(inc))
        dep_cache[path] = tuple(sorted(set(found)))
        return dep_cache[path]

    hdr_hashes = {src: hashlib.sha1(b"".join(p.read_bytes() for p in deps(src))).hexdigest() for src in srcs}   # (sequential: deps() memoises)
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    compile_flags = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                     "-Xcompiler", "-fPIC", "-c"]

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        stamp = objdir / (src.stem + ".sha1")
        # content hash of the unit and every header it includes, transitively (mtimes are not trusted: a concurrent or interrupted build
        # could leave a newer object compiled from older headers)
        want = hashlib.sha1((hdr_hashes[src] + hashlib.sha1(src.read_bytes()).hexdigest() + " ".join(compile_flags)).encode()).hexdigest()
        if obj.exists() and stamp.exists() and stamp.read_text() == want:
            return obj, ""
        stamp.unlink(missing_ok=True)
        cmd = ["nvcc", *compile_flags, *( ["-Xptxas=-v"] if verbose else []), "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise SgbpoError(f"nvcc failed on {src.name}:\n{res.stderr[-4000:]}")
        stamp.write_text(want)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            print(log)
    link_want = hashlib.sha1("".join((objdir / (s.stem + ".sha1")).read_text() for s in srcs).encode()).hexdigest()
    link_stamp = objdir / "libsgbpo.sha1"
    if not LIB_PATH.exists() or not link_stamp.exists() or link_stamp.read_text() != link_want:
        res = subprocess.run(["nvcc", "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
                              *map(str, objs)], capture_output=True, text=True)
        if res.returncode != 0:
            raise SgbpoError(f"nvcc link failed:\n{res.stderr[-4000:]}")
        link_stamp.write_text(link_want)
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
