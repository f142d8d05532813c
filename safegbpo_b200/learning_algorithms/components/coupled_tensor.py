"""(T, N, ...) storage indexed by a per-environment time index
(reference: src/learning_algorithms/components/coupled_tensor.py:7-70)."""
import torch
from torch import Tensor


class CoupledTensor:
    def __init__(self, *shape, dtype=None):
        self.tensor = torch.zeros(shape, dtype=dtype or torch.get_default_dtype())
        self.helper_index = torch.arange(shape[1])

    def reset(self):
        self.tensor = torch.zeros_like(self.tensor)          # fresh storage: drops the autograd history

    def __getitem__(self, index):
        if isinstance(index, Tensor):
            return self.tensor[index, self.helper_index]
        return self.tensor[index]

    def __setitem__(self, key, value):
        if isinstance(key, Tensor):
            self.tensor[key, self.helper_index] = value
        else:
            self.tensor[key] = value

    def nonzero(self):
        return self.tensor.nonzero()

    def count_nonzero(self):
        return self.tensor.count_nonzero()

    def sum(self, dim=None):
        return self.tensor.sum(dim=dim)

    def __mul__(self, other):
        return self.tensor * other

    def __add__(self, other):
        return self.tensor + other

    def __sub__(self, other):
        return self.tensor - other
