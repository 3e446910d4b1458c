"""Drop-in for ``lapy.shapedna`` (reference lapy/shapedna.py): ``compute_shapedna`` (:96-166) runs
on the device; the O(k) post-processing helpers ``normalize_ev`` / ``reweight_ev`` /
``compute_distance`` (:169-296) are plain NumPy on the k eigenvalues and use the
geometry's own ``area()`` / ``volume()`` / ``boundary_tria()`` (a real ``lapy`` mesh has all of
them; the minimal meshes of :mod:`lapy_b200.mesh` provide ``area()``)."""

from __future__ import annotations

import logging

import numpy as np

from .solver import Solver

logger = logging.getLogger(__name__)


def compute_shapedna(geom, k: int = 50, lump: bool = False, aniso=None, aniso_smooth: int = 10,
                     use_cholmod: bool = False) -> dict:  # fmt: skip
    fem = Solver(geom, lump=lump, aniso=aniso, aniso_smooth=aniso_smooth, use_cholmod=use_cholmod)
    evals, evecs = fem.eigs(k=k)
    ev = {"Refine": 0, "Degree": 1}
    if type(geom).__name__ == "TriaMesh":
        ev["Dimension"] = 2
    elif type(geom).__name__ == "TetMesh":
        ev["Dimension"] = 3
    ev["Elements"] = len(geom.t)
    ev["DoF"] = len(geom.v)
    ev["NumEW"] = k
    ev["Eigenvalues"] = evals
    ev["Eigenvectors"] = evecs
    return ev


def _positive(value: float, name: str) -> float:
    if value <= 0:
        raise ValueError(f"{name} must be positive for normalization")
    return value


def _measure(geom, what: str) -> float:
    """area / enclosed volume of the geometry through its own methods (shapedna.py:50-93)."""
    try:
        if what == "area":
            return _positive(geom.area(), "area")
        if type(geom).__name__ == "TriaMesh":
            return _positive(geom.volume(), "volume")
        bnd = geom.boundary_tria()
        bnd.orient_()
        return _positive(bnd.volume(), "boundary volume")
    except AttributeError as e:
        raise NotImplementedError(f"{type(geom).__name__} does not provide the mesh measure needed here ({e}); "
                                  "pass a lapy.TriaMesh / lapy.TetMesh") from e  # fmt: skip


def normalize_ev(geom, evals: np.ndarray, method: str = "geometry") -> np.ndarray:
    """Eigenvalues scaled to unit surface area or unit volume (shapedna.py:169-227).

    ``method``: 'surface' (x area), 'volume' (x volume^(2/3)), 'geometry' (surface for a TriaMesh,
    volume of the oriented boundary for a TetMesh).
    """
    kind = type(geom).__name__
    if method == "surface":
        return evals * _measure(geom, "area")
    if method in ("volume", "geometry"):
        if kind == "TriaMesh":
            return evals * (_measure(geom, "area") if method == "geometry" else _measure(geom, "volume") ** (2.0 / 3.0))
        if kind == "TetMesh":
            return evals * _measure(geom, "volume") ** (2.0 / 3.0)
        raise ValueError(f"Unsupported geometry type for {method} normalization")
    raise ValueError(f"Unknown normalization method: {method}")


def reweight_ev(evals: np.ndarray) -> np.ndarray:
    """``evals[i] / (i + 1)`` (shapedna.py:230-255)."""
    return evals / np.arange(1, len(evals) + 1)


def compute_distance(ev1: np.ndarray, ev2: np.ndarray, dist: str = "euc"):
    """Euclidean distance of two ShapeDNA descriptors (shapedna.py:258-296); any other ``dist`` logs a
    warning and raises ``ValueError`` like the reference."""
    if dist == "euc":
        u, v = np.asarray(ev1, dtype=float), np.asarray(ev2, dtype=float)
        if u.ndim != 1 or v.ndim != 1:
            raise ValueError("Input vector should be 1-D.")
        return float(np.sqrt(np.dot(u - v, u - v)))
    logger.warning("Only Euclidean distance is currently implemented; received %s", dist)
    raise ValueError(f"Distance metric {dist} is not implemented.")
