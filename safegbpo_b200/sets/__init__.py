"""Batched convex sets with the reference's API (src/sets/__init__.py)."""
from .interface import ConvexSet
from .zonotope import Zonotope
from .box import Box
from .axis_aligned_box import AxisAlignedBox
from .ball import Ball

__all__ = ["ConvexSet", "Zonotope", "Box", "AxisAlignedBox", "Ball"]
