"""tcgen05/TMEM critic kernel (csrc/sgbpo_mlp_tc.cu) against a float64 torch evaluation of the same network
(reference: value_function.py:38-66).  Tolerance: 2e-5 absolute on O(1) values -- the fp32 bar of the path;
the fp16 hi/lo split keeps 22 mantissa bits, the FMA-tile kernel (SGBPO_ROWS_TC=0) is held to the same bound."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch


def _run(vf, x):
    from safegbpo_b200 import _lib
    from safegbpo_b200.learning_algorithms.fused_rollout import PackedMlp
    pack = PackedMlp(vf.model)
    out = torch.empty(x.shape[0], dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().sgbpo_mlp_rows(_lib.dtype_code(x.dtype), x.shape[0], C.byref(pack.desc), C.c_void_p(x.data_ptr()),
                                         C.c_void_p(out.data_ptr()), _lib.stream_ptr()), "sgbpo_mlp_rows")
    torch.cuda.synchronize()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("in_dim,depth", [(3, 2), (5, 3), (8, 3), (23, 2)])
def test_tc_rows_match_float64_reference(in_dim, depth):
    from safegbpo_b200.learning_algorithms.components.value_function import ValueFunction
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    torch.set_default_device("cuda:0")
    try:
        torch.manual_seed(11 + in_dim)
        vf = ValueFunction(in_dim, net_arch=[128] * depth)
        with torch.no_grad():                       # trained-looking parameters: non-trivial LayerNorm affine, larger weights
            for m in vf.model:
                if isinstance(m, torch.nn.LayerNorm):
                    m.weight.uniform_(0.5, 2.0)
                    m.bias.uniform_(-0.5, 0.5)
                if isinstance(m, torch.nn.Linear):
                    m.weight.mul_(2.0)
        ref_net = ValueFunction(in_dim, net_arch=[128] * depth).double()
        ref_net.load_state_dict({k: v.double() for k, v in vf.state_dict().items()})
        for M in (1, 127, 128, 129, 4099, 40000):
            x = torch.randn(M, in_dim) * 3
            with torch.no_grad():
                ref = ref_net(x.double()).squeeze(1)
            for tc in ("1", "0"):
                os.environ["SGBPO_ROWS_TC"] = tc
                got = _run(vf, x)
                np.testing.assert_allclose(got.double().cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-5,
                                           err_msg=f"M={M} tensor_core={tc}")
    finally:
        os.environ.pop("SGBPO_ROWS_TC", None)
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)


def test_tensor_core_kernel_is_tcgen05():
    """The shipped library contains the Blackwell tensor-core path: UTCHMMA (tcgen05.mma) and LDTM (tcgen05.ld)."""
    from safegbpo_b200 import _lib
    obj = _lib.PKG / "build" / "sgbpo_mlp_tc.o"
    target = obj if obj.exists() else _lib.LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", str(target)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)], ids=["f64", "f32"])
def test_rows_vjp_matches_torch_autograd(dtype, tol):
    """sgbpo_mlp_rows_vjp (input gradient of the critic on the terminal rows, shac.py:137-147) against torch autograd."""
    from safegbpo_b200 import _lib
    from safegbpo_b200.learning_algorithms.components.value_function import ValueFunction
    from safegbpo_b200.learning_algorithms.fused_rollout import PackedMlp, mlp_vjp_fits
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    torch.set_default_device("cuda:0")
    try:
        torch.manual_seed(9)
        for in_dim, width, depth, M in ((5, 128, 3, 8192), (3, 64, 2, 77), (8, 32, 1, 5), (5, 128, 2, 4099)):
            if not mlp_vjp_fits(dtype, in_dim, width, depth):
                continue
            vf = ValueFunction(in_dim, net_arch=[width] * depth)
            with torch.no_grad():
                for m in vf.model:
                    if isinstance(m, torch.nn.LayerNorm):
                        m.weight.uniform_(0.5, 2.0)
                        m.bias.uniform_(-0.5, 0.5)
            x = (torch.randn(M, in_dim) * 2).requires_grad_(True)
            w = torch.randn(M)
            w[M // 3: 2 * M // 3] = 0.0          # zero-weight rows (unused terminal slots): skipped tiles must still write zeros
            ref = torch.autograd.grad((vf(x).squeeze(1) * w).sum(), x)[0]
            pack = PackedMlp(vf.model)
            gx = torch.empty(M, in_dim)
            _lib.check(_lib.lib().sgbpo_mlp_rows_vjp(_lib.dtype_code(dtype), M, C.byref(pack.desc), C.c_void_p(x.detach().data_ptr()),
                                                     C.c_void_p(w.data_ptr()), C.c_void_p(gx.data_ptr()), _lib.stream_ptr()),
                       "sgbpo_mlp_rows_vjp")
            torch.cuda.synchronize()
            scale = float(ref.abs().max()) + 1e-30
            np.testing.assert_allclose(gx.cpu().numpy() / scale, ref.cpu().numpy() / scale, rtol=10 * tol, atol=tol,
                                       err_msg=f"in={in_dim} W={width} depth={depth} M={M}")
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)
