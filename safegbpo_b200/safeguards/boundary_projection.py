"""Boundary projection (reference: src/safeguards/boundary_projection.py; program P1)."""
from .. import _lib
from .interfaces.safeguard import Safeguard


class BoundaryProjectionSafeguard(Safeguard):
    """Projects an unsafe action onto the closest safe action."""
    SG_KIND = _lib.SG_BOUNDARY_PROJECTION

    def __init__(self, env, **kwargs):
        kwargs.pop("regularisation_coefficient", None)
        Safeguard.__init__(self, env, **kwargs)
