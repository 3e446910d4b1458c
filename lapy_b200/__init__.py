"""lapy_b200 - B200-native (sm_100a) backend for the FEM hot path of Deep-MI/LaPy.

Drop-ins for ``lapy.Solver`` and the callers on the path (``heat.diffusion``,
``diffgeo.compute_geodesic_f`` / gradient / divergence, ``shapedna.compute_shapedna``) plus the
ShapeDNA post-processing and ``.ev`` file format on either side of it (``shapedna.normalize_ev`` …,
``io.read_ev`` / ``write_ev``); everything else of LaPy (mesh toolboxes, other IO, plotting) is out
of scope and keeps working with these objects.
Importing the package does not need a GPU; the first device call does (no CPU fallback).
"""

from . import diffgeo, heat, io, mesh, shapedna  # noqa: F401
from .mesh import TetMesh, TriaMesh  # noqa: F401
from .solver import Solver  # noqa: F401

__version__ = "0.1.0"
