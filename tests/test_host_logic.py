"""CPU-side host logic of the fused engine that needs no GPU: optimizer eligibility, buffer fast-reset hook,
launch-shape bookkeeping mirrored in python."""
import torch

from safegbpo_b200 import optim as sgoptim
from safegbpo_b200.learning_algorithms.components.coupled_buffer import CoupledBuffer
from safegbpo_b200.learning_algorithms.fused_rollout import mlp_rows_fits, net_smem_floats, rollout_fits


def test_fused_optimizer_eligibility():
    p = [torch.nn.Parameter(torch.zeros(3))]
    assert not sgoptim.supported(torch.optim.Adam(p, lr=1e-3))                       # CPU parameters
    assert not sgoptim.supported(torch.optim.SGD(p, lr=1e-3))
    assert not sgoptim.supported(torch.optim.Adam(p, lr=1e-3, weight_decay=0.1))
    assert not sgoptim.supported(torch.optim.AdamW(p, lr=1e-3))


def test_buffer_fast_reset_hook_is_optional_and_falls_back():
    buf = CoupledBuffer(5, 3, 2, True, 1)
    obs0, v0 = torch.arange(6.0).view(3, 2), torch.tensor([1.0, 2.0, 3.0])
    buf.reset(obs0, v0)
    assert torch.equal(buf.observations.tensor[0], obs0) and torch.equal(buf.values.tensor[0], v0)
    calls = []

    def hook(obs=None, value=None):
        calls.append((obs, value))
        return False                      # engine declines -> generic reset must run

    buf.fast_reset = hook
    buf.add(obs0 + 1, torch.zeros(3), torch.zeros(3, dtype=torch.bool), v0 + 1, torch.zeros(3, 1))
    buf.reset()
    assert len(calls) == 1 and calls[0] == (None, None)
    assert torch.equal(buf.observations.tensor[0], obs0 + 1) and torch.equal(buf.values.tensor[0], v0 + 1)
    assert int(buf.t.sum()) == 0
    buf.fast_reset = lambda obs=None, value=None: True      # engine took care of it: storage untouched
    before = buf.observations.tensor
    buf.reset()
    assert buf.observations.tensor is before


def test_shared_memory_fit_predicates():
    # policy [128,128] of the headline workload and the [128]x3 critic fit one SM; fp64 [128]x3 critic does not
    assert rollout_fits(torch.float32, 5, 128, 2, 1) and rollout_fits(torch.float64, 5, 128, 2, 1)
    assert mlp_rows_fits(torch.float32, 5, 128, 3) and not mlp_rows_fits(torch.float64, 5, 128, 3)
    assert net_smem_floats(5, 128, 2, 1) % 8 == 0
