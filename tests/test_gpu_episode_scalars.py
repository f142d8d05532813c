"""sgbpo_episode_stats / sgbpo_shac_loss (csrc/sgbpo_critic.cu) against the torch expressions they replace in the graph
bodies of fused_rollout.py (shac.py:119-120, 131-142; safeguard.py:68-80)."""
import ctypes as C

import pytest
import torch

from safegbpo_b200 import _lib
from safegbpo_b200._lib import check, dtype_code, stream_ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("H,N", [(64, 4096), (7, 40), (1, 1)])
def test_episode_scalars_match_torch(dtype, H, N):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(H * 1000 + N)
    rewards = torch.randn(H + 2, N, generator=g, device=dev, dtype=dtype)
    values = torch.randn(H + 2, N, generator=g, device=dev, dtype=dtype)
    flags = torch.randint(0, 1 << 10, (H, N), generator=g, device=dev, dtype=torch.int32)
    p = lambda t: C.c_void_p(t.data_ptr())
    lib = _lib.lib()
    for bad_mask in (_lib.FL_NONFINITE | _lib.FL_INFEASIBLE, 1 << 20):          # second mask: no flag word has the bit
        scal = torch.zeros(2, dtype=torch.float64, device=dev)
        n_interv = torch.zeros((), dtype=torch.int64, device=dev)
        bad = torch.zeros((), dtype=torch.int32, device=dev)
        for _ in range(2):      # twice: the ticket counter must be left reset
            check(lib.sgbpo_episode_stats(dtype_code(dtype), rewards.numel(), p(rewards), flags.numel(), p(flags), _lib.FL_INTERVENED,
                                          bad_mask, p(scal), p(n_interv), p(bad), stream_ptr()), "sgbpo_episode_stats")
        want_bad = bool(((flags & bad_mask) != 0).any())
        assert int(n_interv) == int(((flags & _lib.FL_INTERVENED) != 0).sum())
        assert int(bad) == int(want_bad) and float(scal[1]) == float(want_bad)
        torch.testing.assert_close(scal[0], rewards.double().sum(), rtol=1e-12 if dtype == torch.float64 else 1e-6, atol=1e-9)
    disc = torch.rand(H + 2, generator=g, device=dev, dtype=dtype)
    disc[0] = 0.0
    term_int = (torch.rand(H + 2, generator=g, device=dev) < 0.3).to(torch.int32)
    norm = torch.full((1,), float(N * H), device=dev, dtype=dtype)
    loss = torch.zeros((), device=dev, dtype=dtype)
    for _ in range(2):
        check(lib.sgbpo_shac_loss(dtype_code(dtype), H + 2, N, p(disc), p(term_int), p(values), p(rewards), p(norm), p(loss),
                                  stream_ptr()), "sgbpo_shac_loss")
    picked = torch.where((term_int != 0).unsqueeze(1), values, rewards).double()
    want = -(disc.double().unsqueeze(1) * picked).sum() / float(N * H)
    torch.testing.assert_close(loss.double(), want, rtol=1e-12 if dtype == torch.float64 else 2e-6, atol=1e-9)
