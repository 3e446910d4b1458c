"""Oracle: eigensolve, Poisson, heat diffusion, heat-method geodesics (TEST INFRASTRUCTURE).

Restates lapy/solver.py:667-716 (``eigs``), :718-889 (``poisson``), lapy/heat.py:114-232
(``diffusion``), lapy/diffgeo.py:116-165 (``compute_geodesic_f``) and
lapy/shapedna.py:96-166 (``compute_shapedna``) on top of the same SciPy SuperLU / ARPACK calls.
This is also the CPU baseline arm timed by ``bench.py`` (sequential SuperLU + ARPACK, 1 core).
"""

from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import LinearOperator, eigsh, splu

from . import diffgeo as _dg
from . import fem as _fem


def eigs(a, b, k=10, sigma=-0.01):
    """k eigenpairs of (A, B) nearest sigma: splu(A - sigma B) + ARPACK mode 3 - solver.py:703-716."""
    lu = splu(a - sigma * b)
    op_inv = LinearOperator(matvec=lu.solve, shape=a.shape, dtype=a.dtype)
    return eigsh(a, k, b, sigma=sigma, OPinv=op_inv)


def poisson(a, m, h=0.0, dtup=(), ntup=()):
    """Solve A x = M (h - n) - A d with Dirichlet elimination - solver.py:771-889."""
    dim = a.shape[0]
    dtype = a.dtype
    if np.isscalar(h):
        h = np.full((dim, 1), h, dtype=dtype)
    else:
        h = np.asarray(h, dtype=dtype)
        if h.ndim == 1:
            h = h[:, None]
    n_rhs = h.shape[1]
    squeeze = n_rhs == 1

    def cols(idx, val):
        idx, val = np.asarray(idx), np.asarray(val)
        return sparse.csc_matrix(
            (np.tile(val, n_rhs), (np.tile(idx, n_rhs), np.repeat(np.arange(n_rhs), len(idx)))),
            (dim, n_rhs),
            dtype=dtype,
        )

    didx = np.asarray(dtup[0]) if dtup else np.zeros(0, int)
    nvec = cols(*ntup) if ntup else 0
    rhs = m.astype(dtype, copy=False) * (h - nvec)
    if len(didx):
        dvec = cols(*dtup)
        rhs = rhs - a * dvec
        keep = np.full(dim, True)
        keep[didx] = False
        rhs = rhs[keep]
        a_ff = a.tocsc()[:, keep].tocsr()[keep, :].tocsc()
    else:
        a_ff = a
    x = splu(a_ff).solve(np.asarray(rhs).astype(dtype))
    if len(didx):
        full = np.zeros((dim, n_rhs), dtype=dtype)
        full[keep, :] = x
        full[didx, :] = dvec[didx, :].toarray()
        x = full
    return np.squeeze(np.array(x)) if squeeze else np.array(x)


def _seed_sets(vids, nv):
    """heat.py:191-202: nesting level decides single vs. multiple seed sets."""
    if isinstance(vids, list) and len(vids) > 0 and isinstance(vids[0], (list, np.ndarray)):
        single, sets = False, [np.asarray(s, dtype=int).ravel() for s in vids]
    else:
        single, sets = True, [np.asarray(vids, dtype=int).ravel()]
    for s in sets:
        if np.any(s < 0) or np.any(s >= nv):
            raise ValueError("vids contains out-of-range vertex indices")
    return single, sets


def diffusion(mesh, vids, m=1.0, aniso=None):
    """Backward-Euler heat from seed vertices: (M_lumped + t A) u = b0 - heat.py:186-232.

    ``aniso`` here is the ready (u1,u2,aniso_mat) triple (curvature is out of scope)."""
    nv = len(mesh.v)
    single, sets = _seed_sets(vids, nv)
    a, b = _fem.fem(mesh, lump=True, aniso=aniso)
    t = m * mesh.avg_edge_length() ** 2
    hmat = b + t * a
    b0 = np.zeros((nv, len(sets)))
    for c, s in enumerate(sets):
        b0[s, c] = 1.0
    u = splu(hmat).solve(b0.astype(np.float64))
    return u[:, 0] if single else u


def geodesic_f(mesh, f):
    """Function with unit gradient field of f (heat method step 2+3) - diffgeo.py:144-165."""
    f = np.asarray(f)
    g = _dg.gradient(mesh, f)
    a, _ = _fem.fem(mesh, lump=True)
    eye = sparse.eye(a.shape[0], dtype=a.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        if f.ndim == 1:
            gn = np.nan_to_num(g / np.sqrt((g**2).sum(1))[:, None])
        else:
            gn = np.nan_to_num(g / np.sqrt((g**2).sum(-1))[:, :, None])
    vf = poisson(a, eye, _dg.divergence(mesh, gn))
    vf -= vf.min() if f.ndim == 1 else vf.min(axis=0)
    return vf


def shapedna(mesh, k=50, lump=False, aniso=None):
    """shapedna.py:146-166."""
    a, b = _fem.fem(mesh, lump=lump, aniso=aniso)
    evals, evecs = eigs(a, b, k=k)
    return {
        "Refine": 0,
        "Degree": 1,
        "Dimension": 2 if type(mesh).__name__ == "TriaMesh" else 3,
        "Elements": len(mesh.t),
        "DoF": len(mesh.v),
        "NumEW": k,
        "Eigenvalues": evals,
        "Eigenvectors": evecs,
    }
