"""Drop-ins for the hot-path functions of ``lapy.diffgeo`` (reference lapy/diffgeo.py):
``compute_gradient`` (:27-56), ``compute_divergence`` (:59-113), ``compute_geodesic_f`` (:116-165)
and their tria / tet variants, with the reference's 1-D / 2-D shape conventions, plus the loop
callers that re-use the same kernels: ``tria_compute_divergence2`` (:390-469),
``tria_compute_rotated_f`` (:472-520), ``tria_mean_curvature_flow`` (:523-607) and
``tria_spherical_project`` (:634-843)."""

from __future__ import annotations

import logging
import math

import numpy as np

from . import _lib
from .solver import Solver, _device_mesh

logger = logging.getLogger(__name__)


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
    identity mass (``fem.mass = eye`` in the reference: the right-hand side is the divergence
    itself); the minimum is shifted to 0 per column.

    Gradient, normalisation (with the reference's ``nan_to_num``) and divergence run back to back on
    the device (``lb_unit_gradient_divergence``): no (n_elements, 3) field crosses PCIe, the mesh is
    uploaded once."""
    _check(geom, None)
    vfunc = np.asarray(vfunc)
    scalar_input = vfunc.ndim == 1
    ctx = _lib.default_context()
    mesh = _device_mesh(geom, ctx)
    if vfunc.shape[0] != mesh.nv:
        raise ValueError("vfunc needs one value per vertex")
    divf = _lib.unit_gradient_divergence(ctx, mesh, vfunc.reshape(mesh.nv, -1))
    fem = Solver(geom, lump=True, use_cholmod=use_cholmod, ctx=ctx, _mesh=mesh)
    if divf.shape[0] != fem._shape0():
        raise ValueError("h should be either scalar or array with first dim matching A")
    vf, info = _lib.solve(ctx, fem._device("a"), 1.0, None, 0.0, divf, project_nullspace=True)
    compute_geodesic_f.last_info = info
    vf = vf[:, 0] if scalar_input else vf
    if scalar_input:
        vf -= vf.min()
    else:
        vf -= vf.min(axis=0)
    return vf


def tria_compute_geodesic_f(tria, vfunc, use_cholmod: bool = False):
    _check(tria, "TriaMesh")
    return compute_geodesic_f(tria, vfunc, use_cholmod)


def tria_compute_divergence2(tria, tfunc):
    """Integrated divergence in flux form: sum of <tfunc, e_ij x n> over the 1-ring (diffgeo.py:390-469);
    equal to :func:`tria_compute_divergence` up to rounding."""
    _check(tria, "TriaMesh")
    tfunc = np.asarray(tfunc)
    ctx = _lib.default_context()
    mesh = _device_mesh(tria, ctx)
    if tfunc.shape[0] != mesh.nt or tfunc.shape[-1] != 3:
        raise ValueError("tfunc needs one 3-vector per triangle")
    d = _lib.divergence(ctx, mesh, tfunc.reshape(mesh.nt, -1, 3), flux=True)
    return d[:, 0] if tfunc.ndim == 2 else d


def tria_compute_rotated_f(tria, vfunc, use_cholmod: bool = False):
    """Function whose gradient is the gradient of ``vfunc`` rotated by 90 degrees about the triangle
    normals (diffgeo.py:472-520): Poisson problem for the divergence of the rotated gradient with
    vertex 0 pinned to 0 and an identity mass."""
    _check(tria, "TriaMesh")
    vfunc = np.asarray(vfunc)
    scalar_input = vfunc.ndim == 1
    ctx = _lib.default_context()
    mesh = _device_mesh(tria, ctx)
    gradf = _lib.gradient(ctx, mesh, vfunc.reshape(mesh.nv, -1))  # (nt, nf, 3)
    tn = tria.tria_normals()
    rot = np.cross(tn[:, np.newaxis, :], gradf)
    divf = _lib.divergence(ctx, mesh, rot)
    fem = Solver(tria, lump=True, use_cholmod=use_cholmod, ctx=ctx, _mesh=mesh)
    if divf.shape[0] != fem._shape0():
        raise ValueError("h should be either scalar or array with first dim matching A")
    vf, info = _lib.solve(ctx, fem._device("a"), 1.0, None, 0.0, divf, fix_idx=np.array([0]),
                          fix_val=np.array([0.0]))  # fmt: skip
    tria_compute_rotated_f.last_info = info
    return vf[:, 0] if scalar_input else vf


def tria_mean_curvature_flow(tria, max_iter: int = 30, stop_eps: float = 1e-13, step: float = 1.0,
                             use_cholmod: bool = False):  # fmt: skip
    """Conformalised mean curvature flow (Kazhdan 2012; diffgeo.py:523-607): the stiffness matrix of the
    normalised input mesh stays fixed, every step re-assembles the lumped mass for the current
    vertices and solves ``(M + step*A) v_new = M v`` for the three coordinates, then re-normalises
    (centroid at the origin, unit area).  Returns a new ``TriaMesh``.

    On the device: one mesh upload; per step ``lb_mesh_update_vertices`` + a mass-only assembly +
    one 3-column solve (AMG-preconditioned block CG: the operator is stiffness dominated)."""
    _check(tria, "TriaMesh")
    cls = type(tria)
    trianorm = cls(tria.v, tria.t)
    trianorm.normalize_()
    ctx = _lib.default_context()
    mesh = _lib.DeviceMesh(ctx, trianorm.v, trianorm.t)
    a_dev, _b0 = _lib.assemble(ctx, mesh, _lib.FEM_TRIA, True)
    for it in range(max_iter):
        vlast = trianorm.v
        if it > 0:
            mesh.update_vertices(trianorm.v)
        _none, mass = _lib.assemble(ctx, mesh, _lib.FEM_TRIA_MASS, True, want_a=False)
        mass_v = _lib.spmm(ctx, mass, np.asarray(vlast, dtype=np.float64))
        vnew, info = _lib.solve(ctx, a_dev, float(step), mass, 1.0, mass_v)
        trianorm.v = vnew
        trianorm.normalize_()
        dv = trianorm.v - vlast
        diff = np.trace(np.square(np.matmul(np.transpose(dv), _lib.spmm(ctx, mass, dv))))
        logger.debug("Step %d delta: %g", it + 1, diff)
        if diff < stop_eps:
            logger.info("Converged after %d iterations.", it + 1)
            break
    tria_mean_curvature_flow.last_info = info
    return trianorm


def _unit_vector(v: np.ndarray, name: str) -> np.ndarray:
    norm = np.linalg.norm(v)
    if np.isclose(norm, 0.0):
        raise ValueError(f"{name} is degenerate (zero length)")
    return v / norm


def _flipped_area(mesh) -> float:
    """Area of the triangles whose normal points towards the origin (meaningful on a centred sphere)."""
    p1, p2, p3 = (mesh.v[mesh.t[:, c], :] for c in range(3))
    cr = np.cross(p2 - p1, p3 - p1)
    inward = np.sum(p1 * cr, axis=1) < 0
    return np.sum((0.5 * np.sqrt(np.sum(cr * cr, axis=1)))[inward])


def tria_spherical_project(tria, flow_iter: int = 3, debug: bool = False):
    """Spectral embedding by the first three non-constant eigenfunctions, a few mean curvature flow
    steps, projection onto the sphere of radius 100 (diffgeo.py:634-843), with the reference's
    orientation rules (FreeSurfer axes) and sanity checks (same ``ValueError`` messages)."""
    _check(tria, "TriaMesh")
    if not tria.is_closed():
        raise ValueError("Error: Can only project closed meshes!")
    fem = Solver(tria, lump=False)
    evals, evecs = fem.eigs(k=4)
    if debug:
        from .io import write_ev

        write_ev({"Eigenvalues": evals, "Eigenvectors": evecs, "Creator": "spherically_project.py", "Refine": 0,
                  "Degree": 1, "Dimension": 2, "Elements": tria.t.shape[0], "DoF": evecs.shape[0], "NumEW": 4},
                 "debug.ev")  # fmt: skip

    def extremes(ev):
        return np.mean(tria.v[ev > 0.5 * np.max(ev), :], 0), np.mean(tria.v[ev < 0.5 * np.min(ev), :], 0)

    ev1, ev2, ev3 = evecs[:, 1], evecs[:, 2], evecs[:, 3]
    (cmax1, cmin1), (cmax2, cmin2), (cmax3, cmin3) = extremes(ev1), extremes(ev2), extremes(ev3)
    # eigenfunction 1 is trusted to run front to back (axis 1 = y for brains in FreeSurfer space)
    l11, l21, l31 = abs(cmax1[1] - cmin1[1]), abs(cmax2[1] - cmin2[1]), abs(cmax3[1] - cmin3[1])
    if l11 < l21 or l11 < l31:
        logger.error("Direction 1 should be anterior - posterior (%g, %g, %g)", l11, l21, l31)
        raise ValueError("Direction 1 should be anterior - posterior")
    v1 = _unit_vector(cmax1 - cmin1, "direction 1")
    if cmax1[1] < cmin1[1]:
        ev1 = -1 * ev1
    l1 = abs(cmax1[1] - cmin1[1])
    # eigenfunctions 2 / 3: superior - inferior (z) and right - left (x), swapped if 3 fits z better
    if abs(cmax2[2] - cmin2[2]) < abs(cmax3[2] - cmin3[2]):
        ev2, ev3 = ev3, ev2
        cmax2, cmax3 = cmax3, cmax2
        cmin2, cmin3 = cmin3, cmin2
    if abs(cmax3[0] - cmin3[0]) < abs(cmax2[0] - cmin2[0]):
        logger.warning("WARNING: direction 3 wants to swap with 2, but cannot")
    v2 = _unit_vector(cmax2 - cmin2, "direction 2")
    if cmax2[2] < cmin2[2]:
        ev2 = -1 * ev2
    l2 = abs(cmax2[2] - cmin2[2])
    v3 = _unit_vector(cmax3 - cmin3, "direction 3")
    if cmax3[0] < cmin3[0]:
        ev3 = -1 * ev3
    l3 = abs(cmax3[0] - cmin3[0])
    spatvol = abs(np.dot(v1, np.cross(v2, v3)))
    mvol = tria.volume()
    logger.info("box %g, %g, %g volume: %g, coverage %g", l1, l2, l3, l1 * l2 * l3, l1 * l2 * l3 / mvol)
    # map every eigenfunction to -1 .. 0 .. +1 keeping the zero level set fixed
    scaled = []
    for ev in (ev1, ev2, ev3):
        ev = np.array(ev)
        lo, hi = np.amin(ev), np.amax(ev)
        ev[ev < 0] /= -lo
        ev[ev > 0] /= hi
        scaled.append(ev)
    vn = np.empty(tria.v.shape)
    vn[:, 0], vn[:, 1], vn[:, 2] = scaled[2], scaled[0], scaled[1]
    cls = type(tria)
    if flow_iter > 0:
        vn = tria_mean_curvature_flow(cls(vn, tria.t), max_iter=flow_iter).v
    dist = np.sqrt(np.sum(vn * vn, axis=1))
    trianew = cls(100 * (vn / dist[:, np.newaxis]), tria.t)
    svol = trianew.area() / (4.0 * math.pi * 10000)
    flippedarea = _flipped_area(trianew) / (4.0 * math.pi * 10000)
    if flippedarea > 0.95:
        logger.error("global normal flip detected: %g", flippedarea)
        raise ValueError("global normal flip")
    if svol < 0.99:
        logger.error("sphere area fraction below threshold .99 > %g", svol)
        raise ValueError("sphere area fraction should be above .99")
    if flippedarea > 0.0008:
        logger.error("flipped area fraction too high (>0.0008): %g", flippedarea)
        raise ValueError("flipped area fraction should be below .0008")
    if spatvol < 0.6:
        logger.error("spat vol (orthogonality) below threshold 0.6 > %g", spatvol)
        raise ValueError("spat vol (orthogonality) should be above .6")
    return trianew
