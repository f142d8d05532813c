"""ConvexSet ABC (reference: src/sets/interface/convex_set.py:9-82)."""
from abc import ABC, abstractmethod


class ConvexSet(ABC):
    def __init__(self, batch_dim: int, dim: int):
        self.batch_dim, self.dim = batch_dim, dim

    @abstractmethod
    def sample(self): ...

    @abstractmethod
    def contains(self, other): ...

    @abstractmethod
    def intersects(self, other): ...

    def draw(self, ax=None, **kwargs):
        raise NotImplementedError("plotting is outside the hot path (SURVEY §2.1)")
