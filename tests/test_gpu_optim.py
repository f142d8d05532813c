"""Fused clip_grad_norm_ + Adam launch (csrc/sgbpo_optim.cu) against torch's own implementation
(reference: src/learning_algorithms/shac.py:148-149)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)], ids=["f64", "f32"])
@pytest.mark.parametrize("scale", [0.01, 30.0], ids=["unclipped", "clipped"])
def test_clip_adam_matches_torch(dtype, tol, scale):
    from safegbpo_b200 import optim as sgoptim
    from safegbpo_b200.learning_algorithms.components.policy import Policy
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    torch.set_default_device("cuda:0")
    try:
        torch.manual_seed(5)
        a, b = Policy(5, 2, net_arch=[64, 64]), Policy(5, 2, net_arch=[64, 64])
        b.load_state_dict(a.state_dict())
        oa = torch.optim.Adam(a.parameters(), lr=3e-3, capturable=True)
        ob = torch.optim.Adam(b.parameters(), lr=3e-3, capturable=True)
        assert sgoptim.supported(ob)
        for it in range(5):
            grads = [torch.randn_like(p) * scale for p in a.parameters()]
            for p, q, g in zip(a.parameters(), b.parameters(), grads):
                p.grad, q.grad = g.clone(), g.clone()
            ref_norm = torch.nn.utils.clip_grad_norm_(a.parameters(), 1.0)
            oa.step()
            got_norm = sgoptim.clip_adam_step(ob, 1.0)
            np.testing.assert_allclose(float(got_norm), float(ref_norm), rtol=10 * tol)
            for (name, p), q in zip(a.named_parameters(), b.parameters()):
                np.testing.assert_allclose(q.grad.cpu().numpy(), p.grad.cpu().numpy(), rtol=10 * tol, atol=tol, err_msg=f"grad {name} it {it}")
                np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().cpu().numpy(), rtol=10 * tol, atol=tol, err_msg=f"{name} it {it}")
                sa, sb = oa.state[p], ob.state[q]
                assert float(sa["step"]) == float(sb["step"]) == it + 1
                np.testing.assert_allclose(sb["exp_avg_sq"].cpu().numpy(), sa["exp_avg_sq"].cpu().numpy(), rtol=10 * tol, atol=tol * tol)
                np.testing.assert_allclose(sb["exp_avg"].cpu().numpy(), sa["exp_avg"].cpu().numpy(), rtol=10 * tol, atol=tol)
        # the optimizer state stays loadable by torch
        oc = torch.optim.Adam(b.parameters(), lr=3e-3, capturable=True)
        oc.load_state_dict(ob.state_dict())
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)


def test_value_function_inference_path_matches_torch():
    """ValueFunction.forward under no_grad on CUDA goes through the batched MLP kernel (one launch); with autograd
    enabled it stays the torch module.  Both must agree."""
    from safegbpo_b200.learning_algorithms.components.value_function import ValueFunction
    prev = torch.get_default_dtype()
    torch.set_default_device("cuda:0")
    try:
        for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 1e-11)):
            torch.set_default_dtype(dtype)
            torch.manual_seed(3)
            for arch in ([128, 128, 128], [64, 64], [48, 48]):      # tensor-core kernel, FMA tiles, unsupported width (torch)
                vf = ValueFunction(5, net_arch=arch)
                x = torch.randn(777, 5) * 2
                ref = vf(x)
                assert ref.requires_grad
                with torch.no_grad():
                    got = vf(x)
                assert got.shape == ref.shape and not got.requires_grad
                np.testing.assert_allclose(got.cpu().numpy(), ref.detach().cpu().numpy(), rtol=tol, atol=tol)
    finally:
        torch.set_default_device("cpu")
        torch.set_default_dtype(prev)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 3e-5)], ids=["f64", "f32"])
def test_bias_elu_layernorm_rows_match_torch_autograd(dtype, tol):
    """Fused hidden-layer tail of the critic fit (csrc/sgbpo_critic.cu) against torch: h = LayerNorm(ELU(z + b)),
    gradients w.r.t. z, bias, gamma, beta (value_function.py:38-48 / shac.py:154-184)."""
    from safegbpo_b200.learning_algorithms.shac import _BiasEluLayerNormRows
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    for W, M in ((32, 7), (64, 1000), (128, 4099)):
        z = (torch.randn(M, W, dtype=dtype, device=dev) * 2).requires_grad_(True)
        b = torch.randn(W, dtype=dtype, device=dev).requires_grad_(True)
        g = (torch.rand(W, dtype=dtype, device=dev) + 0.5).requires_grad_(True)
        be = torch.randn(W, dtype=dtype, device=dev).requires_grad_(True)
        cot = torch.randn(M, W, dtype=dtype, device=dev)
        ref = torch.nn.functional.layer_norm(torch.nn.functional.elu(z + b), (W,), g, be, 1e-5)
        ref_grads = torch.autograd.grad((ref * cot).sum(), (z, b, g, be))
        got = _BiasEluLayerNormRows.apply(z, b, g, be)
        got_grads = torch.autograd.grad((got * cot).sum(), (z, b, g, be))
        np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=tol, atol=tol)
        for name, a, r in zip(("dz", "dbias", "dgamma", "dbeta"), got_grads, ref_grads):
            scale = float(r.abs().max()) + 1e-30
            np.testing.assert_allclose(a.cpu().numpy() / scale, r.cpu().numpy() / scale, rtol=10 * tol, atol=10 * tol,
                                       err_msg=f"{name} W={W} M={M}")
