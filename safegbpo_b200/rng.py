"""Single choke point for every random draw of the rollout, in the reference's call order
(policy rsample -> ray-mask directions -> noise sample -> reset draws).

Default: torch's global generator on the default device, exactly the calls the reference makes
(`torch.rand(N, 1, dim)`, `Normal.rsample`).  Tests install a replay source so a CPU-recorded stream of
the reference / oracle can be fed to the CUDA path for seed-identical parity."""
from __future__ import annotations

from typing import Callable, Optional

import torch

_source: Optional[Callable] = None


def set_source(fn: Optional[Callable]):
    """fn(kind, shape) -> Tensor on the default device/dtype, kind in {"rand", "randn"}; None restores torch."""
    global _source
    _source = fn


def rand(*shape) -> torch.Tensor:
    if _source is not None:
        return _source("rand", tuple(shape))
    return torch.rand(*shape)


def randn(*shape) -> torch.Tensor:
    if _source is not None:
        return _source("randn", tuple(shape))
    return torch.randn(*shape)
