"""Drop-in for ``lapy.shapedna.compute_shapedna`` (reference lapy/shapedna.py:96-166).  The
O(k) post-processing helpers (normalize_ev, reweight_ev, compute_distance) stay with the
reference (out of scope)."""

from __future__ import annotations

from .solver import Solver


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
