"""Ball (obstacles of the Navigate* tasks; reference: src/sets/ball.py:12-152)."""
import torch
from torch import Tensor

from .. import rng
from .interface import ConvexSet


class Ball(ConvexSet):
    def __init__(self, center: Tensor, radius: Tensor):
        super().__init__(*center.shape)
        self.center, self.radius = center, radius

    def sample(self) -> Tensor:
        direction = rng.randn(self.batch_dim, self.dim)
        direction = direction / torch.linalg.vector_norm(direction, dim=1, keepdim=True)
        r = self.radius.unsqueeze(1) * rng.rand(self.batch_dim, 1) ** (1.0 / self.dim)
        return self.center + direction * r

    def contains(self, other) -> Tensor:
        if isinstance(other, Tensor):
            return torch.linalg.vector_norm(other - self.center, dim=1) <= self.radius
        if isinstance(other, Ball):
            return torch.linalg.vector_norm(other.center - self.center, dim=1) + other.radius <= self.radius
        raise NotImplementedError(f"Containment check not implemented for {type(other)}")

    def intersects(self, other) -> Tensor:
        from .box import Box
        if isinstance(other, Ball):
            return torch.linalg.vector_norm(other.center - self.center, dim=1) <= self.radius + other.radius
        if isinstance(other, Box):
            local = torch.sum((self.center - other.center).unsqueeze(2) * other.edge_dir, dim=1)
            closest = torch.clamp(local, -other.edge_len, other.edge_len)
            return torch.linalg.vector_norm(local - closest, dim=1) <= self.radius
        raise NotImplementedError(f"Intersection check not implemented for {type(other)}")
