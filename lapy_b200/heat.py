"""Drop-in for ``lapy.heat``: ``diffusion`` (reference lapy/heat.py:114-232) runs on the device;
``diagonal`` / ``kernel`` (:19-111) are the reference's small dense products on the
eigenpairs ``Solver.eigs`` returned, evaluated on the host arrays the caller holds."""

from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .solver import Solver

logger = logging.getLogger(__name__)


def diagonal(t, x, evecs: np.ndarray, evals: np.ndarray, n: int) -> np.ndarray:
    """Heat kernel diagonal ``K(t, x, x) = sum_j exp(-lambda_j t) phi_j(x)^2`` over the first ``n``
    eigenpairs (heat.py:19-59): rows = vertices ``x``, columns = times ``t``."""
    if n > evecs.shape[1] or n > evals.shape[0]:
        raise ValueError("n exceeds the number of available eigenpairs")
    sq = evecs[x, 0:n] * evecs[x, 0:n]
    return np.matmul(sq, np.exp(-np.matmul(evals[0:n], t)))


def kernel(t, vfix: int, evecs: np.ndarray, evals: np.ndarray, n: int) -> np.ndarray:
    """Heat kernel ``K_t(p, vfix) = sum_j exp(-lambda_j t) phi_j(p) phi_j(vfix)`` from all vertices to
    one fixed vertex over the first ``n`` eigenpairs (heat.py:62-111): rows = vertices, columns = times.

    The reference's expression multiplies the (n, n_times) exponentials by the (n,) row
    ``evecs[vfix]`` without a trailing axis and only broadcasts for n_times == n; this is the formula
    of its docstring.  ``evals`` may be (k,) or the (k, 1) column the reference documents, ``t`` a
    scalar or any array of times."""
    if n > evecs.shape[1] or n > evals.shape[0]:
        raise ValueError("n exceeds the number of available eigenpairs")
    lam = np.asarray(evals).reshape(-1)[0:n]
    times = np.asarray(t, dtype=float).reshape(-1)
    weights = np.exp(-lam[:, None] * times[None, :]) * evecs[vfix, 0:n][:, None]
    return np.matmul(evecs[:, 0:n], weights)


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
    if np.asarray(geometry.v).dtype == np.float64:
        # device kernel over the stiffness pattern (same edge set as triu(adj_sym, 1), lapy/tria_mesh.py:735-748,
        # lapy/tet_mesh.py:182-195); the host version costs seconds at millions of vertices.  float32 meshes
        # keep the geometry's own (float32, pairwise-summed) arithmetic so that t is bit-identical to the reference's.
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
