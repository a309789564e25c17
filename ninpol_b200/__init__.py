"""ninpol_b200 — B200-native (sm_100a) implementation of ninpol's nodal-interpolation hot path.

Drop-in surface: `Interpolator` (load_mesh, interpolate, supported_methods) and its `grid`; see
INTEGRATION.md.  Usage mirrors the reference: `import ninpol_b200 as ninpol`.
"""
from .interpolator import Interpolator
from .grid import Grid
from . import meshgen, dist

__all__ = ["Interpolator", "Grid", "meshgen", "dist"]
__version__ = "0.1.0"
