"""Drop-in for the hot-path part of ``lapy.heat`` (reference lapy/heat.py:114-232)."""

from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .solver import Solver

logger = logging.getLogger(__name__)


def diffusion(geometry, vids, m: float = 1.0, aniso=None, use_cholmod: bool = False, *, tol: float = 0.0):
    """Heat diffusion from seed vertices by one backward-Euler step ``(B + t A) u = b0``,
    ``t = m * avg_edge_length**2``, lumped mass.  Same ``vids`` nesting rules, error and
    return shapes as the reference: (n,) for a single seed set, (n, n_cases) for a list of sets.
    All seed sets are solved together by one block CG on the GPU (heat.py:210-227 solves them
    through one factorisation)."""
    nv = len(geometry.v)
    if isinstance(vids, list) and len(vids) > 0 and isinstance(vids[0], (list, np.ndarray)):
        scalar_input = False
        vids_list = [np.asarray(v, dtype=int).ravel() for v in vids]
    else:
        scalar_input = True
        vids_list = [np.asarray(vids, dtype=int).ravel()]
    for v in vids_list:
        if np.any(v < 0) or np.any(v >= nv):
            raise ValueError("vids contains out-of-range vertex indices")
    fem = Solver(geometry, lump=True, aniso=aniso)
    if np.asarray(geometry.v).dtype == np.float64 and nv >= 100000:
        # device kernel over the stiffness pattern (same edge set as triu(adj_sym, 1)); the host
        # version costs seconds at millions of vertices.  float32 meshes keep the geometry's own
        # (float32) arithmetic so that t is bit-identical to the reference's.
        t = m * _lib.avg_edge_length(fem._ctx, fem._mesh, fem._device("a")) ** 2
    else:
        t = m * geometry.avg_edge_length() ** 2
    n = fem._shape0()
    b0 = np.zeros((n, len(vids_list)))
    for k, v in enumerate(vids_list):
        b0[v, k] = 1.0
    logger.info("Solver: Jacobi/AMG-preconditioned block CG on the GPU")
    x, info = _lib.solve(fem._ctx, fem._device("a"), float(t), fem._device("b"), 1.0, b0, tol=tol)
    diffusion.last_info = info
    if scalar_input:
        return x[:, 0]
    return x
