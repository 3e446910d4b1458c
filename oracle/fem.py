"""Oracle: linear-FEM stiffness / mass assembly (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates lapy/solver.py:105-194 (``_fem_tria``), :196-308 (``_fem_tria_aniso``), :310-377
(``fem_tria_mass``) and :379-533 (``_fem_tetra``) in structure-of-arrays form.  All local
arithmetic runs in the dtype of ``v`` (float32 meshes stay float32 until the final widening,
SURVEY.md §0.4); operation order is that of the reference so results are bit-identical:
``dot(x,y) = (x0*y0 + x1*y1) + x2*y2`` and ``cross`` unfused (SURVEY.md §7 "Bit-faithful").
"""

from __future__ import annotations

import sys

import numpy as np
from scipy import sparse

EPS = sys.float_info.epsilon

# COO slot order of one triangle / tet (solver.py:171-175, :472-497): (row corner, col corner)
TRIA_SLOTS = ((0, 1), (1, 0), (1, 2), (2, 1), (2, 0), (0, 2), (0, 0), (1, 1), (2, 2))
TET_SLOTS = (
    (0, 1), (1, 0), (1, 2), (2, 1), (2, 0), (0, 2), (0, 3), (3, 0),
    (1, 3), (3, 1), (2, 3), (3, 2), (0, 0), (1, 1), (2, 2), (3, 3),
)  # fmt: skip


def _corner(v, t, c):
    p = v[t[:, c], :]
    return p[:, 0], p[:, 1], p[:, 2]


def _sub(a, b):
    return a[0] - b[0], a[1] - b[1], a[2] - b[2]


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _clamp(vol):
    """solver.py:158-159 / :357-358 / :436-437: degenerate elements get 1e-4 * mean."""
    vol_mean = max(0.0001 * np.mean(vol), EPS)
    vol = vol.copy()
    vol[vol < EPS] = vol_mean
    return vol


def _coo(t, slots, cols):
    i = np.column_stack([t[:, r] for r, _ in slots]).reshape(-1)
    j = np.column_stack([t[:, c] for _, c in slots]).reshape(-1)
    dat = np.column_stack(cols).reshape(-1)
    return dat, i, j


def _csc(dat, i, j, dtype):
    return sparse.csc_matrix((dat.astype(dtype, copy=False), (i, j)), dtype=dtype)


def _lumped(t, b, dtype):
    k = t.shape[1]
    i = np.column_stack([t[:, c] for c in range(k)]).reshape(-1)
    dat = np.column_stack([b] * k).reshape(-1)
    return _csc(dat, i, i, dtype)


def tria_local(v, t, aniso=None):
    """Per-triangle quantities: (a12, a23, a31, vol) with vol = 4*area, clamped.

    solver.py:145-165; with ``aniso=(u1,u2,aniso_mat)`` the projected variant :259-280.
    """
    p1, p2, p3 = _corner(v, t, 0), _corner(v, t, 1), _corner(v, t, 2)
    e_c, e_a, e_b = _sub(p2, p1), _sub(p3, p2), _sub(p1, p3)  # v2mv1, v3mv2, v1mv3
    cr = _cross(e_a, e_b)
    vol = _clamp(2 * np.sqrt(_dot(cr, cr)))
    if aniso is None:
        a12 = _dot(e_a, e_b) / vol
        a23 = _dot(e_b, e_c) / vol
        a31 = _dot(e_c, e_a) / vol
    else:
        u1, u2, am = aniso
        u1 = (u1[:, 0], u1[:, 1], u1[:, 2])
        u2 = (u2[:, 0], u2[:, 1], u2[:, 2])

        def proj(e):
            return _dot(u1, e), _dot(u2, e)

        def adot(x, y):
            return x[0] * am[:, 0] * y[0] + x[1] * am[:, 1] * y[1]

        ua, ub, uc = proj(e_a), proj(e_b), proj(e_c)
        a12 = adot(ua, ub) / vol
        a23 = adot(ub, uc) / vol
        a31 = adot(uc, ua) / vol
    return a12, a23, a31, vol


def _tria_matrices(t, a12, a23, a31, vol, lump, dtype):
    a11 = -a12 - a31
    a22 = -a12 - a23
    a33 = -a31 - a23
    dat, i, j = _coo(t, TRIA_SLOTS, (a12, a12, a23, a23, a31, a31, a11, a22, a33))
    a = _csc(dat, i, j, dtype)
    if lump:
        b = _lumped(t, vol / 12, dtype)
    else:
        bii, bij = vol / 24, vol / 48
        dat, i, j = _coo(t, TRIA_SLOTS, (bij,) * 6 + (bii,) * 3)
        b = _csc(dat, i, j, dtype)
    return a, b


def fem_tria(v, t, lump=False, dtype=np.float64):
    """(A, B) as scipy CSC - solver.py:105-194."""
    dtype = np.dtype(dtype)
    return _tria_matrices(t, *tria_local(v, t), lump, dtype)


def fem_tria_aniso(v, t, u1, u2, aniso_mat, lump=False, dtype=np.float64):
    """(A, B) as scipy CSC - solver.py:196-308."""
    dtype = np.dtype(dtype)
    return _tria_matrices(t, *tria_local(v, t, (u1, u2, aniso_mat)), lump, dtype)


def fem_tria_mass(v, t, lump=False, dtype=np.float64):
    """Mass matrix only, vol = area - solver.py:310-377."""
    dtype = np.dtype(dtype)
    p1, p2, p3 = _corner(v, t, 0), _corner(v, t, 1), _corner(v, t, 2)
    cr = _cross(_sub(p3, p2), _sub(p1, p3))
    vol = _clamp(0.5 * np.sqrt(_dot(cr, cr)))
    if lump:
        return _lumped(t, vol / 3, dtype)
    bii, bij = vol / 6, vol / 12
    dat, i, j = _coo(t, TRIA_SLOTS, (bij,) * 6 + (bii,) * 3)
    return _csc(dat, i, j, dtype)


def tet_local(v, t):
    """Per-tet off-diagonals BEFORE the final /6 and vol = 6*volume (clamped).

    solver.py:418-465.  Returns ((a12,a13,a14,a23,a24,a34), vol).
    """
    p1, p2, p3, p4 = (_corner(v, t, c) for c in range(4))
    e1, e2, e3 = _sub(p2, p1), _sub(p3, p2), _sub(p1, p3)
    e4, e5, e6 = _sub(p4, p1), _sub(p4, p2), _sub(p4, p3)
    vol = _clamp(np.abs(_dot(e4, _cross(e1, e3))))
    e11, e22, e33 = _dot(e1, e1), _dot(e2, e2), _dot(e3, e3)
    e44, e55, e66 = _dot(e4, e4), _dot(e5, e5), _dot(e6, e6)
    e12, e13, e14, e15 = _dot(e1, e2), _dot(e1, e3), _dot(e1, e4), _dot(e1, e5)
    e23, e25, e26 = _dot(e2, e3), _dot(e2, e5), _dot(e2, e6)
    e34, e36 = _dot(e3, e4), _dot(e3, e6)
    a12 = (-e36 * e26 + e23 * e66) / vol
    a13 = (-e15 * e25 + e12 * e55) / vol
    a14 = (e23 * e26 - e36 * e22) / vol
    a23 = (-e14 * e34 + e13 * e44) / vol
    a24 = (e13 * e34 - e14 * e33) / vol
    a34 = (-e14 * e13 + e11 * e34) / vol
    return (a12, a13, a14, a23, a24, a34), vol


def fem_tetra(v, t, lump=False, dtype=np.float64):
    """(A, B) as scipy CSC - solver.py:379-533."""
    dtype = np.dtype(dtype)
    (a12, a13, a14, a23, a24, a34), vol = tet_local(v, t)
    a11 = -a12 - a13 - a14
    a22 = -a12 - a23 - a24
    a33 = -a13 - a23 - a34
    a44 = -a14 - a24 - a34
    cols = (a12, a12, a23, a23, a13, a13, a14, a14, a24, a24, a34, a34, a11, a22, a33, a44)
    dat, i, j = _coo(t, TET_SLOTS, cols)
    a = _csc(dat / 6.0, i, j, dtype)
    if lump:
        b = _lumped(t, vol / 24.0, dtype)
    else:
        bii, bij = vol / 60.0, vol / 120.0
        dat, i, j = _coo(t, TET_SLOTS, (bij,) * 12 + (bii,) * 4)
        b = _csc(dat, i, j, dtype)
    return a, b


def fem(mesh, lump=False, aniso=None, dtype=np.float64):
    """Dispatch like ``Solver.__init__`` (solver.py:72-99); ``aniso=(u1,u2,aniso_mat)``."""
    name = type(mesh).__name__
    if name == "TriaMesh":
        if aniso is not None:
            return fem_tria_aniso(mesh.v, mesh.t, *aniso, lump=lump, dtype=dtype)
        return fem_tria(mesh.v, mesh.t, lump, dtype)
    if name == "TetMesh":
        return fem_tetra(mesh.v, mesh.t, lump, dtype)
    raise ValueError('Geometry type "' + name + '" unknown')
