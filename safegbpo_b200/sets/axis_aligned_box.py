"""AxisAlignedBox: diagonal generator, cached min/max (reference: src/sets/axis_aligned_box.py:8-31)."""
import torch
from torch import Tensor

from .box import Box


class AxisAlignedBox(Box):
    def __init__(self, center: Tensor, generator: Tensor):
        diag = torch.diagonal(generator, dim1=-2, dim2=-1)
        if not torch.equal(generator, torch.diag_embed(diag)):
            raise AssertionError("Generator must be diagonal for AxisAlignedBox.")
        super().__init__(center, generator)
        self.min = center - diag
        self.max = center + diag
