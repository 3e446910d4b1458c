"""lapy_b200 - B200-native (sm_100a) backend for the FEM hot path of Deep-MI/LaPy.

Drop-ins for ``lapy.Solver`` and the callers on the path (``heat.diffusion``,
``diffgeo.compute_geodesic_f`` / gradient / divergence, ``shapedna.compute_shapedna``); everything
else of LaPy (mesh toolboxes, IO, plotting) is out of scope and keeps working with these objects.
Importing the package does not need a GPU; the first device call does (no CPU fallback).
"""

from . import diffgeo, heat, mesh, shapedna  # noqa: F401
from .mesh import TetMesh, TriaMesh  # noqa: F401
from .solver import Solver  # noqa: F401

__version__ = "0.1.0"
