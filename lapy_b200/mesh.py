"""Minimal mesh containers and the synthetic benchmark meshes.

The FEM hot path of the reference only reads ``geometry.v`` (n,3 float32/float64) and
``geometry.t`` (m,3|4 integer) and dispatches on ``type(geometry).__name__``
(reference lapy/solver.py:72, :95).  The reference's full ``TriaMesh`` / ``TetMesh``
toolboxes (lapy/tria_mesh.py, lapy/tet_mesh.py) are out of scope; a real
``lapy.TriaMesh`` works with :class:`lapy_b200.Solver` unchanged.  The two containers here
exist so that tests, ``bench.py`` and ``smoke()`` run on a box where the reference is not
installed.  They carry the same names on purpose (name based duck typing).

Generators (host side, NumPy only, deterministic, no RNG unless a seed is given):

* :func:`icosphere` - BASELINE.json config 2 / BASELINE.md §5.2.  Vertex and triangle order
  follow the reference's 1-to-4 ``refine_`` (lapy/tria_mesh.py:1452-1487: parents first, then
  one midpoint per edge in (min,max)-lexicographic edge order; children ``[a,ab,ca]``,
  ``[b,bc,ab]``, ``[c,ca,bc]``, ``[ab,bc,ca]``), so eigenvectors are comparable to goldens.
* :func:`cube_tets` - BASELINE.json config 3 / SURVEY.md §8d.3 (6-tet cell pattern of
  ``data/cubeTetra.vtk``).
* :func:`perturbed_sphere` - BASELINE.json config 5 (BrainPrint-like batch surfaces).
"""

from __future__ import annotations

import numpy as np

__all__ = ["TriaMesh", "TetMesh", "icosphere", "cube_tets", "perturbed_sphere"]


def _unique_edges(t: np.ndarray, nv: int):
    """Undirected edges of an element array as sorted (lo, hi) pairs, lexicographic."""
    k = t.shape[1]
    pairs = [(a, b) for a in range(k) for b in range(a + 1, k)]
    lo = np.concatenate([np.minimum(t[:, a], t[:, b]) for a, b in pairs]).astype(np.int64)
    hi = np.concatenate([np.maximum(t[:, a], t[:, b]) for a, b in pairs]).astype(np.int64)
    key = np.unique(lo * nv + hi)
    return key // nv, key % nv


class _Mesh:
    def __init__(self, v, t, width):
        self.v = np.array(v)
        self.t = np.array(t)
        if self.v.ndim != 2 or self.t.ndim != 2:
            raise ValueError("v and t must be 2-D arrays")
        if self.v.shape[0] < self.v.shape[1]:
            self.v = self.v.T
        if self.t.shape[1] != width and self.t.shape[0] == width:
            self.t = self.t.T
        if self.t.size == 0 or self.v.size == 0:
            raise ValueError("empty mesh")
        if self.t.shape[1] != width:
            raise ValueError(f"elements should have {width} vertices")
        if self.v.shape[1] == 2:
            self.v = np.column_stack([self.v, np.zeros(self.v.shape[0])])
        if self.v.shape[1] != 3:
            raise ValueError("vertices should have 2 or 3 coordinates")
        if np.max(self.t) >= self.v.shape[0]:
            raise ValueError("Max index exceeds number of vertices")

    def avg_edge_length(self) -> float:
        """Mean length of the unique undirected edges (reference lapy/tria_mesh.py:735-748,
        lapy/tet_mesh.py:182-195: ``triu(adj_sym, 1)`` rows/cols, same arithmetic in the
        dtype of ``v``)."""
        lo, hi = _unique_edges(self.t, self.v.shape[0])
        d = self.v[lo, :] - self.v[hi, :]
        return np.sqrt((d**2).sum(1)).mean()


class TriaMesh(_Mesh):
    """Triangle mesh: ``v`` (n,3), ``t`` (m,3)."""

    def __init__(self, v, t):
        super().__init__(v, t, 3)

    def tria_areas(self) -> np.ndarray:
        """Triangle areas by Heron's formula from the three edge lengths, like the reference
        (lapy/tria_mesh.py:636-658), so ``area()`` agrees with ``lapy.TriaMesh.area()``."""
        p0, p1, p2 = (self.v[self.t[:, c], :] for c in range(3))
        a, b, c = (np.sqrt(np.sum(e * e, axis=1)) for e in (p1 - p0, p2 - p1, p0 - p2))
        half = 0.5 * (a + b + c)
        return np.sqrt(half * (half - a) * (half - b) * (half - c))

    def area(self) -> float:
        """Total surface area (lapy/tria_mesh.py:660-669)."""
        return np.sum(self.tria_areas())

    def boundary_vertices(self) -> np.ndarray:
        """Sorted indices of vertices on edges that belong to exactly one triangle."""
        nv = self.v.shape[0]
        t = self.t.astype(np.int64)
        lo = np.concatenate([np.minimum(t[:, a], t[:, b]) for a, b in ((0, 1), (1, 2), (2, 0))])
        hi = np.concatenate([np.maximum(t[:, a], t[:, b]) for a, b in ((0, 1), (1, 2), (2, 0))])
        key, cnt = np.unique(lo * nv + hi, return_counts=True)
        b = key[cnt == 1]
        return np.unique(np.concatenate([b // nv, b % nv]))


class TetMesh(_Mesh):
    """Tetrahedral mesh: ``v`` (n,3), ``t`` (m,4)."""

    def __init__(self, v, t):
        super().__init__(v, t, 4)


# The 12 vertices / 20 faces of the reference's data/icosahedron.off as its OFF reader
# delivers them (float32 of the 6-decimal text, radius 2).
_ICO_V = np.array(
    [
        [0.0, 0.0, 2.0],
        [1.788854, 0.0, 0.894427],
        [0.552786, 1.701302, 0.894427],
        [-1.447214, 1.051462, 0.894427],
        [-1.447214, -1.051462, 0.894427],
        [0.552786, -1.701302, 0.894427],
        [1.447214, 1.051462, -0.894427],
        [-0.552786, 1.701302, -0.894427],
        [-1.788854, 0.0, -0.894427],
        [-0.552786, -1.701302, -0.894427],
        [1.447214, -1.051462, -0.894427],
        [0.0, 0.0, -2.0],
    ],
    dtype=np.float32,
)
_ICO_T = np.array(
    [
        [2, 0, 1], [3, 0, 2], [4, 0, 3], [5, 0, 4], [1, 0, 5],
        [2, 1, 6], [7, 2, 6], [3, 2, 7], [8, 3, 7], [4, 3, 8],
        [9, 4, 8], [5, 4, 9], [10, 5, 9], [6, 1, 10], [1, 5, 10],
        [6, 11, 7], [7, 11, 8], [8, 11, 9], [9, 11, 10], [10, 11, 6],
    ],
    dtype=np.int64,
)  # fmt: skip


def _refine_once(v: np.ndarray, t: np.ndarray):
    nv = v.shape[0]
    lo, hi = _unique_edges(t, nv)
    mid = 0.5 * (v[lo, :] + v[hi, :])
    vnew = np.append(v, mid, axis=0)
    keys = lo * nv + hi

    def eid(a, b):
        k = np.minimum(a, b).astype(np.int64) * nv + np.maximum(a, b)
        return nv + np.searchsorted(keys, k)

    a, b, c = t[:, 0], t[:, 1], t[:, 2]
    ab, bc, ca = eid(a, b), eid(b, c), eid(c, a)
    tnew = np.concatenate(
        (
            np.column_stack((a, ab, ca)),
            np.column_stack((b, bc, ab)),
            np.column_stack((c, ca, bc)),
            np.column_stack((ab, bc, ca)),
        ),
        axis=1,
    ).reshape(-1, 3)
    return vnew, tnew


def icosphere(level: int) -> TriaMesh:
    """Unit icosphere: ``level`` x (1-to-4 refine, cast to float64, project to the sphere).

    Level 9: 2,621,442 vertices / 5,242,880 triangles (BASELINE.json config 2).
    Level 0 returns the raw float32 radius-2 icosahedron like the reference's reader.
    """
    v, t = _ICO_V.copy(), _ICO_T.copy()
    for _ in range(level):
        v, t = _refine_once(v, t)
        v = v.astype(np.float64)
        v /= np.linalg.norm(v, axis=1)[:, None]
    return TriaMesh(v, t)


_CELL_TETS = (
    ((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)),
    ((0, 0, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1)),
    ((1, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1)),
    ((0, 0, 1), (0, 1, 0), (1, 1, 1), (0, 1, 1)),
    ((0, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 0)),
    ((1, 1, 0), (0, 0, 1), (0, 1, 0), (1, 1, 1)),
)


def cube_tets(n: int) -> TetMesh:
    """Unit cube with ``n`` vertices per side, index = x + n*y + n*n*z, 6 tets per cell.

    n=121: 1,771,561 vertices / 10,368,000 tets (BASELINE.json config 3).
    """
    g = np.arange(n, dtype=np.float64) / (n - 1)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    v = np.column_stack((x.reshape(-1), y.reshape(-1), z.reshape(-1)))
    c = np.arange(n - 1, dtype=np.int64)
    cz, cy, cx = np.meshgrid(c, c, c, indexing="ij")
    base = (cx + n * cy + n * n * cz).reshape(-1)
    cols = []
    for tet in _CELL_TETS:
        for dx, dy, dz in tet:
            cols.append(base + dx + n * dy + n * n * dz)
    t = np.column_stack(cols).reshape(-1, 4)
    return TetMesh(v, t)


def perturbed_sphere(level: int, seed: int, amp: float = 0.2) -> TriaMesh:
    """Star-shaped smooth radial perturbation of :func:`icosphere`:
    ``r = 1 + amp/8 * sum_j c_j sin(w_j . x + phi_j)``, seeded per mesh."""
    s = icosphere(level)
    rng = np.random.default_rng(seed)
    w = rng.normal(size=(8, 3)) * 2.0
    c = rng.uniform(-1.0, 1.0, size=8)
    phi = rng.uniform(0.0, 2.0 * np.pi, size=8)
    r = 1.0 + (amp / 8.0) * (np.sin(s.v @ w.T + phi) * c).sum(1)
    return TriaMesh(s.v * r[:, None], s.t)
