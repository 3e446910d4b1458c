"""Oracle: per-element gradient and integrated divergence (TEST INFRASTRUCTURE).

Restates lapy/diffgeo.py:222-300 (``tria_compute_gradient``), :303-387
(``tria_compute_divergence``), :846-922 (``tet_compute_gradient``) and :925-1006
(``tet_compute_divergence``).  1-D inputs follow the reference's 1-D code path, 2-D inputs
(n, F) / (T, F, 3) its batched path; the vertex scatter uses the same SciPy calls the
reference uses (1-column ``csc_matrix`` for 1-D, CSR scatter matrix product for 2-D) so the
per-vertex summation order is the reference's.
"""

from __future__ import annotations

import sys

import numpy as np
from scipy import sparse

EPS = sys.float_info.epsilon


def _edges_tria(v, t):
    v0, v1, v2 = v[t[:, 0], :], v[t[:, 1], :], v[t[:, 2], :]
    return v2 - v1, v0 - v2, v1 - v0  # e0, e1, e2


def tria_gradient(v, t, f):
    """grad f per triangle: (T,3) for f (n,), (T,F,3) for f (n,F) - diffgeo.py:258-300."""
    f = np.asarray(f)
    e0, e1, e2 = _edges_tria(v, t)
    n = np.cross(e2, -e1)
    ln = np.sqrt(np.sum(n * n, axis=1))
    ln[ln < EPS] = 1
    lni = np.divide(1.0, ln)[:, None]
    n = n * lni
    if f.ndim == 1:
        s = f[t[:, 0], None] * e0 + f[t[:, 1], None] * e1 + f[t[:, 2], None] * e2
        return lni * np.cross(n, s)
    f0, f1, f2 = f[t[:, 0], :], f[t[:, 1], :], f[t[:, 2], :]
    s = (
        f0[:, :, None] * e0[:, None, :]
        + f1[:, :, None] * e1[:, None, :]
        + f2[:, :, None] * e2[:, None, :]
    )
    return lni[:, None, :] * np.cross(n[:, None, :], s)


def _scatter(t, xs, nv, scale, dtype):
    """Element-corner values -> vertex sums, reference summation order (see module doc)."""
    k = t.shape[1]
    if xs[0].ndim == 1:
        i = np.column_stack([t[:, c] for c in range(k)]).reshape(-1)
        j = np.zeros(k * len(t), dtype=int)
        dat = np.column_stack(xs).reshape(-1)
        return np.squeeze(np.asarray(scale * sparse.csc_matrix((dat, (i, j))).todense(), dtype=dtype))
    rows = np.concatenate([t[:, c] for c in range(k)])
    s = sparse.csr_matrix((np.ones(k * len(t)), (rows, np.arange(k * len(t)))), shape=(nv, k * len(t)))
    return scale * s.dot(np.vstack(xs))


def tria_divergence(v, t, x):
    """Integrated divergence at vertices of a per-triangle field - diffgeo.py:331-387."""
    x = np.asarray(x)
    e0, e1, e2 = _edges_tria(v, t)
    n = np.cross(e2, -e1)
    ln = np.sqrt(np.sum(n * n, axis=1))
    ln[ln < EPS] = 1
    cot0 = (e2 * (-e1)).sum(1) / ln
    cot1 = (e0 * (-e2)).sum(1) / ln
    cot2 = (e1 * (-e0)).sum(1) / ln
    c0, c1, c2 = cot0[:, None] * e0, cot1[:, None] * e1, cot2[:, None] * e2
    if x.ndim == 2:
        xs = [((c2 - c1) * x).sum(1), ((c0 - c2) * x).sum(1), ((c1 - c0) * x).sum(1)]
    else:
        xs = [
            ((c2 - c1)[:, None, :] * x).sum(-1),
            ((c0 - c2)[:, None, :] * x).sum(-1),
            ((c1 - c0)[:, None, :] * x).sum(-1),
        ]
    return _scatter(t, xs, v.shape[0], 0.5, x.dtype)


def tet_gradient(v, t, f):
    """grad f per tet - diffgeo.py:877-922."""
    f = np.asarray(f)
    v0, v1, v2, v3 = (v[t[:, c], :] for c in range(4))
    e0, e2, e3, e4, e5 = v1 - v0, v0 - v2, v3 - v0, v3 - v1, v3 - v2
    vol = np.abs(np.sum(e3 * np.cross(e0, e2), axis=1))
    vol[vol < EPS] = 1
    voli = np.divide(1.0, vol)[:, None]
    g1, g2, g3 = np.cross(e2, e5), np.cross(e3, e4), np.cross(-e2, e0)
    if f.ndim == 1:
        d1 = f[t[:, 1], None] - f[t[:, 0], None]
        d2 = f[t[:, 2], None] - f[t[:, 0], None]
        d3 = f[t[:, 3], None] - f[t[:, 0], None]
        return voli * (d1 * g1 + d2 * g2 + d3 * g3)
    d1 = f[t[:, 1], :] - f[t[:, 0], :]
    d2 = f[t[:, 2], :] - f[t[:, 0], :]
    d3 = f[t[:, 3], :] - f[t[:, 0], :]
    s = d1[:, :, None] * g1[:, None, :] + d2[:, :, None] * g2[:, None, :] + d3[:, :, None] * g3[:, None, :]
    return voli[:, None, :] * s


def tet_divergence(v, t, x):
    """Integrated divergence at vertices of a per-tet field - diffgeo.py:950-1006."""
    x = np.asarray(x)
    v0, v1, v2, v3 = (v[t[:, c], :] for c in range(4))
    e0, e1, e2, e3, e4 = v1 - v0, v2 - v1, v2 - v0, v3 - v0, v3 - v1
    ns = [np.cross(e1, e4), np.cross(e3, e2), np.cross(e0, e3), np.cross(e2, e0)]
    if x.ndim == 2:
        xs = [(n * x).sum(1) for n in ns]
        return -_scatter(t, xs, v.shape[0], 1.0 / 6.0, x.dtype)
    xs = [(n[:, None, :] * x).sum(-1) for n in ns]
    return _scatter(t, xs, v.shape[0], -(1.0 / 6.0), x.dtype)


def gradient(mesh, f):
    if type(mesh).__name__ == "TriaMesh":
        return tria_gradient(mesh.v, mesh.t, f)
    if type(mesh).__name__ == "TetMesh":
        return tet_gradient(mesh.v, mesh.t, f)
    raise ValueError('Geometry type "' + type(mesh).__name__ + '" unknown')


def divergence(mesh, x):
    if type(mesh).__name__ == "TriaMesh":
        return tria_divergence(mesh.v, mesh.t, x)
    if type(mesh).__name__ == "TetMesh":
        return tet_divergence(mesh.v, mesh.t, x)
    raise ValueError('Geometry type "' + type(mesh).__name__ + '" unknown')
