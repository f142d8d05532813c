"""Ray mask (reference: src/safeguards/ray_mask.py; programs P2/P3/P4 + radial map)."""
import torch

from .. import _lib, rng
from .interfaces.safeguard import Safeguard


class RayMaskSafeguard(Safeguard):
    """Maps actions radially towards a safe centre."""
    SG_KIND = _lib.SG_RAY_MASK

    def __init__(self, env, linear_projection: bool, zonotopic_approximation: bool, passthrough: bool, **kwargs):
        kwargs.pop("regularisation_coefficient", None)
        Safeguard.__init__(self, env, **kwargs)
        self.linear_projection = linear_projection
        self.zonotopic_approximation = zonotopic_approximation
        self.passthrough = passthrough

    def _rm_flags(self):
        return dict(rm_linear=self.linear_projection, rm_zonotopic=self.zonotopic_approximation,
                    rm_passthrough=self.passthrough)

    def _directions(self):
        """ray_mask.py:184-185: drawn every call inside the safeguard (Q9) when the env is state constrained."""
        if not (self.zonotopic_approximation and self.state_constrained):
            return None
        m = self.action_dim
        d = rng.rand(self.batch_dim, m, 2 * m) * 2 - 1
        return d / torch.linalg.vector_norm(d, dim=1, keepdim=True)
