"""Box = zonotope with dim orthogonal generators (reference: src/sets/box.py:12-178)."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import rng
from .interface import ConvexSet
from .zonotope import Zonotope


class Box(Zonotope):
    def __init__(self, center: Tensor, generator: Tensor):
        if generator.shape[1] != generator.shape[2]:
            raise TypeError("Box needs a square generator")
        super().__init__(center, generator)
        self.edge_len = torch.linalg.vector_norm(self.generator, dim=1)
        if (self.edge_len == 0).any():               # box.py:36-38: complete a degenerate box to a full basis
            u, s, v = torch.svd(self.generator)
            self.edge_dir = u @ v.transpose(1, 2)
        elif generator.isinf().any():
            self.edge_dir = torch.diag_embed(torch.ones(self.batch_dim, self.dim, dtype=center.dtype, device=center.device))
        else:
            self.edge_dir = self.generator / self.edge_len.unsqueeze(1)
        self.edge_len = torch.where(self.edge_len == 0, torch.full_like(self.edge_len, 1.0e-9), self.edge_len)
        self.generator = torch.where(self.generator.isinf(),
                                     torch.full_like(self.generator, torch.finfo(self.generator.dtype).max)
                                     * self.generator.sign(), self.generator)

    def sample(self) -> Tensor:
        return self.center + torch.sum(self.generator * (2 * rng.rand(self.batch_dim, 1, self.dim) - 1), dim=2)

    def _proj(self, center):
        return torch.sum(center.unsqueeze(2) * self.edge_dir, dim=1).abs()

    def contains(self, other) -> Tensor:
        from .ball import Ball
        if isinstance(other, Tensor):
            return (self._proj(other - self.center) <= self.edge_len).all(dim=1)
        if isinstance(other, Ball):
            return (self._proj(other.center - self.center) + other.radius.unsqueeze(1) <= self.edge_len).all(dim=1)
        if isinstance(other, Zonotope):
            supports = (self.edge_dir.unsqueeze(2) * other.generator.unsqueeze(3)).sum(dim=1).abs().sum(dim=1)
            return (self._proj(other.center - self.center) + supports <= self.edge_len).all(dim=1)
        raise NotImplementedError(f"Containment check not implemented for {type(other)}")

    def intersects(self, other: ConvexSet) -> Tensor:
        from .ball import Ball
        if isinstance(other, Ball):
            return other.intersects(self)
        if isinstance(other, Zonotope):               # box.py:123-154 (the Zonotope branch is the gate, Q1)
            center = other.center - self.center
            supports = (self.edge_dir.unsqueeze(2) * other.generator.unsqueeze(3)).sum(dim=1).abs().sum(dim=1)
            if isinstance(other, Box):
                o_dir, o_len = other.edge_dir, other.edge_len
            else:
                o_len = torch.linalg.vector_norm(other.generator, dim=1)
                o_dir = other.generator / o_len.unsqueeze(1)
            o_proj = torch.sum(center.unsqueeze(2) * o_dir, dim=1).abs()
            o_supports = (o_dir.unsqueeze(2) * self.generator.unsqueeze(3)).sum(dim=1).abs().sum(dim=1)
            return (self._proj(center) <= self.edge_len + supports).all(dim=1) & (o_proj <= o_len + o_supports).all(dim=1)
        raise NotImplementedError(f"Intersection check not implemented for {type(other)}")

    def farthest_box_corner(self, point: Tensor) -> Tensor:
        d1 = torch.linalg.vector_norm(self.center.unsqueeze(2) + self.generator - point.unsqueeze(2), dim=1, keepdim=True)
        d2 = torch.linalg.vector_norm(self.center.unsqueeze(2) - self.generator - point.unsqueeze(2), dim=1, keepdim=True)
        return self.center + torch.sum(torch.where(d1 > d2, 1.0, -1.0) * self.generator, dim=2)
