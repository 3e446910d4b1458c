"""Drop-ins for the hot-path functions of ``lapy.diffgeo`` (reference lapy/diffgeo.py):
``compute_gradient`` (:27-56), ``compute_divergence`` (:59-113), ``compute_geodesic_f`` (:116-165)
and their tria / tet variants, with the reference's 1-D / 2-D shape conventions."""

from __future__ import annotations

import numpy as np
from scipy import sparse

from . import _lib
from .solver import Solver, _device_mesh


def _check(geom, want):
    name = type(geom).__name__
    if want is not None and name != want:
        raise ValueError('Geometry type "' + name + '" unknown')
    if name not in ("TriaMesh", "TetMesh"):
        raise ValueError('Geometry type "' + name + '" unknown')


def compute_gradient(geom, vfunc, _want=None):
    """Gradient of a vertex function: (n_elements, 3) for vfunc (n,), (n_elements, F, 3) for (n, F)."""
    _check(geom, _want)
    vfunc = np.asarray(vfunc)
    ctx = _lib.default_context()
    mesh = _device_mesh(geom, ctx)
    if vfunc.shape[0] != mesh.nv:
        raise ValueError("vfunc needs one value per vertex")
    g = _lib.gradient(ctx, mesh, vfunc.reshape(mesh.nv, -1))
    return g[:, 0, :] if vfunc.ndim == 1 else g


def compute_divergence(geom, vfunc, _want=None):
    """Integrated divergence at vertices of a per-element field: (n,) for (n_elements, 3),
    (n, F) for (n_elements, F, 3)."""
    _check(geom, _want)
    vfunc = np.asarray(vfunc)
    ctx = _lib.default_context()
    mesh = _device_mesh(geom, ctx)
    if vfunc.shape[0] != mesh.nt or vfunc.shape[-1] != 3:
        raise ValueError("vfunc needs one 3-vector per element")
    d = _lib.divergence(ctx, mesh, vfunc.reshape(mesh.nt, -1, 3))
    return d[:, 0] if vfunc.ndim == 2 else d


def tria_compute_gradient(tria, vfunc):
    return compute_gradient(tria, vfunc, "TriaMesh")


def tet_compute_gradient(tet, vfunc):
    return compute_gradient(tet, vfunc, "TetMesh")


def tria_compute_divergence(tria, tfunc):
    return compute_divergence(tria, tfunc, "TriaMesh")


def tet_compute_divergence(tet, tfunc):
    return compute_divergence(tet, tfunc, "TetMesh")


def compute_geodesic_f(geom, vfunc, use_cholmod: bool = False):
    """Function with unit gradient along the gradient of ``vfunc`` (heat method, diffgeo.py:144-165):
    normalise grad f, take its integrated divergence and solve the Poisson problem with an
    identity mass; the minimum is shifted to 0 per column."""
    vfunc = np.asarray(vfunc)
    scalar_input = vfunc.ndim == 1
    gradf = compute_gradient(geom, vfunc)
    fem = Solver(geom, lump=True, use_cholmod=use_cholmod)
    fem.mass = sparse.eye(fem.stiffness.shape[0], dtype=fem.stiffness.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        if scalar_input:
            gradnorm = gradf / np.sqrt((gradf**2).sum(1))[:, np.newaxis]
        else:
            gradnorm = gradf / np.sqrt((gradf**2).sum(-1))[:, :, np.newaxis]
    gradnorm = np.nan_to_num(gradnorm)
    divf = compute_divergence(geom, gradnorm)
    vf = fem.poisson(divf)
    if scalar_input:
        vf -= vf.min()
    else:
        vf -= vf.min(axis=0)
    return vf


def tria_compute_geodesic_f(tria, vfunc, use_cholmod: bool = False):
    _check(tria, "TriaMesh")
    return compute_geodesic_f(tria, vfunc, use_cholmod)
