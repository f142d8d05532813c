"""Shared-memory operand views of the tcgen05 path (csrc/sgbpo_tc.cuh): one stored 128 x 128 operand serves as the
K-major operand of  X W^T-style products and as the MN-major operand of  X^T Z-style products (weight gradients).
Checked against float64 torch; tolerance 2e-5 relative to the product scale (fp16 hi/lo split, three MMAs)."""
import ctypes as C

import numpy as np
import pytest
import torch


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_operand_views(mode):
    from safegbpo_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5 + mode)
    x = torch.randn(128, 128, generator=g, dtype=torch.float32, device="cpu").to(dev)      # (explicit: the kernel reads float32)
    z = (torch.randn(128, 128, generator=g, dtype=torch.float32, device="cpu") * 0.5).to(dev)
    d = torch.zeros(128, 128, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().sgbpo_tc_selftest(mode, C.c_void_p(x.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(d.data_ptr()),
                                            _lib.stream_ptr()), "sgbpo_tc_selftest")
    torch.cuda.synchronize()
    xd, zd = x.double(), z.double()
    ref = [xd @ zd.t(), xd @ zd, xd.t() @ zd][mode]
    scale = float(ref.abs().max())
    err = float((d.double() - ref).abs().max()) / scale
    assert err < 2e-5, f"mode {mode}: max error {err:.3e} of the product scale"
