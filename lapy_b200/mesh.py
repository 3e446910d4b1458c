"""Minimal mesh containers and the synthetic benchmark meshes.

The FEM hot path of the reference only reads ``geometry.v`` (n,3 float32/float64) and
``geometry.t`` (m,3|4 integer) and dispatches on ``type(geometry).__name__``
(reference lapy/solver.py:72, :95).  The reference's full ``TriaMesh`` / ``TetMesh``
toolboxes (lapy/tria_mesh.py, lapy/tet_mesh.py: IO, plotting, level sets, refinement ...) are out
of scope; a real ``lapy.TriaMesh`` works with :class:`lapy_b200.Solver` unchanged.
The two containers here carry the same names on purpose (name based duck typing) and provide the
mesh-side pre-steps the callers ON the path need: adjacency
(``adj_sym`` / ``adj_dir``), ``avg_edge_length``, normals, ``curvature`` / ``curvature_tria`` (the
input of the anisotropic operator), ``normalize_`` / ``centroid`` / ``volume`` (mean curvature
flow, spherical projection).  They are host-side NumPy restatements of the reference's formulas
(same operation order, so results agree to rounding), each citing the lines it follows.

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

    # ---- adjacency (lapy/tria_mesh.py:486-534); built on first use, the reference builds them in
    # __init__ (:274-275, 2.5-3.9 s at 2.6M vertices) although the FEM path never reads them
    def _half_edges(self):
        t = self.t.astype(np.int64)
        i = np.column_stack((t[:, 0], t[:, 1], t[:, 2])).reshape(-1)
        j = np.column_stack((t[:, 1], t[:, 2], t[:, 0])).reshape(-1)
        return i, j

    @property
    def adj_dir(self):
        """Directed adjacency: entry (i, j) = number of triangles with the half edge i -> j."""
        if getattr(self, "_adj_dir_t", None) is not self.t:
            from scipy import sparse

            i, j = self._half_edges()
            n = self.v.shape[0]
            self._adj_dir = sparse.csc_matrix((np.ones(i.shape), (i, j)), shape=(n, n))
            self._adj_dir_t = self.t
        return self._adj_dir

    @property
    def adj_sym(self):
        """Symmetric adjacency: entry (i, j) = number of triangles containing the edge {i, j}."""
        if getattr(self, "_adj_sym_t", None) is not self.t:
            from scipy import sparse

            i, j = self._half_edges()
            n = self.v.shape[0]
            self._adj_sym = sparse.csc_matrix((np.ones(2 * i.size), (np.concatenate((i, j)), np.concatenate((j, i)))),
                                              shape=(n, n))  # fmt: skip
            self._adj_sym_t = self.t
        return self._adj_sym

    def is_closed(self) -> bool:
        """No boundary edges (lapy/tria_mesh.py:562-572)."""
        return 1 not in self.adj_sym.data

    def is_manifold(self) -> bool:
        return np.max(self.adj_sym.data) <= 2

    def is_oriented(self) -> bool:
        """Every half edge belongs to one triangle only (lapy/tria_mesh.py:604-615)."""
        return np.max(self.adj_dir.data) == 1

    def edges(self):
        """Unique inner edges ``vids`` (ne, 2), lexicographic with vids[:,0] < vids[:,1], and the two
        triangles ``tids`` (ne, 2): the one holding the half edge vids[:,0] -> vids[:,1] and its
        neighbour across the edge (lapy/tria_mesh.py:1019-1076 without the lil matrices)."""
        if not self.is_oriented():
            raise ValueError("Error: Can only compute edge information for oriented meshes!")
        n = self.v.shape[0]
        i, j = self._half_edges()
        tid = np.repeat(np.arange(self.t.shape[0], dtype=np.int64), 3)
        key = i * n + j
        order = np.argsort(key, kind="stable")
        skey, stid = key[order], tid[order]
        fwd = np.flatnonzero(i < j)
        fwd = fwd[np.argsort(key[fwd], kind="stable")]
        rev = j[fwd] * n + i[fwd]
        pos = np.searchsorted(skey, rev)
        pos_c = np.minimum(pos, skey.size - 1)
        inner = skey[pos_c] == rev  # the opposite half edge exists: not a boundary edge
        fwd, pos_c = fwd[inner], pos_c[inner]
        vids = np.column_stack((i[fwd], j[fwd])).astype(np.int32)
        tids = np.column_stack((tid[fwd], stid[pos_c])).astype(np.int32)
        return vids, tids

    def tria_normals(self) -> np.ndarray:
        """Unit triangle normals (lapy/tria_mesh.py:750-773)."""
        v0, v1, v2 = (self.v[self.t[:, c], :] for c in range(3))
        nrm = np.cross(v1 - v0, v2 - v0)
        ln = np.sqrt(np.sum(nrm * nrm, axis=1))
        ln[ln < np.finfo(float).eps] = 1
        return nrm / ln.reshape(-1, 1)

    def vertex_normals(self) -> np.ndarray:
        """Angle-weighted unit vertex normals (lapy/tria_mesh.py:775-817)."""
        if not self.is_oriented():
            raise ValueError("Error: Vertex normals are meaningless for un-oriented triangle meshes!")
        v0, v1, v2 = (self.v[self.t[:, c], :] for c in range(3))
        e01, e12, e20 = v1 - v0, v2 - v1, v0 - v2
        nrm = np.zeros(self.v.shape)
        np.add.at(nrm, self.t[:, 0], np.cross(e01, -e20))
        np.add.at(nrm, self.t[:, 1], np.cross(e12, -e01))
        np.add.at(nrm, self.t[:, 2], np.cross(e20, -e12))
        ln = np.sqrt(np.sum(nrm * nrm, axis=1))
        ln[ln < np.finfo(float).eps] = 1
        return nrm / ln.reshape(-1, 1)

    def vertex_areas(self) -> np.ndarray:
        """A third of the area of the incident triangles per vertex (lapy/tria_mesh.py:715-733)."""
        v0, v1, v2 = (self.v[self.t[:, c], :] for c in range(3))
        cr = np.cross(v1 - v0, v2 - v0)
        area = 0.5 * np.sqrt(np.sum(cr * cr, axis=1))
        area3 = np.repeat(area[:, np.newaxis], 3, 1)
        return np.bincount(self.t.flatten(), area3.flatten(), minlength=self.v.shape[0]) / 3.0

    def centroid(self):
        """Area-weighted centroid and total area (lapy/tria_mesh.py:990-1017)."""
        v0, v1, v2 = (self.v[self.t[:, c], :] for c in range(3))
        cr = np.cross(v2 - v1, v0 - v2)
        areas = 0.5 * np.sqrt(np.sum(cr * cr, axis=1))
        total = areas.sum()
        centers = (1.0 / 3.0) * (v0 + v1 + v2)
        return np.sum(centers * (areas / total)[:, np.newaxis], axis=0), total

    def normalize_(self) -> None:
        """Centroid to the origin, unit surface area (lapy/tria_mesh.py:1390-1405)."""
        c, area = self.centroid()
        if area <= 0:
            raise ValueError("Mesh surface area must be positive to normalize.")
        self.v = (1.0 / np.sqrt(area)) * (self.v - c)

    def volume(self) -> float:
        """Enclosed volume of a closed, oriented mesh (lapy/tria_mesh.py:671-702)."""
        if not self.is_closed():
            raise ValueError("Mesh must be closed to compute volume.")
        if not self.is_oriented():
            raise ValueError("Mesh must be oriented to compute volume.")
        v0, v1, v2 = (self.v[self.t[:, c], :] for c in range(3))
        return np.sum(np.sum(v0 * np.cross(v1 - v0, v2 - v0), axis=1)) / 6.0

    def map_vfunc_to_tfunc(self, vfunc: np.ndarray) -> np.ndarray:
        """Triangle values = mean of the three vertex values (lapy/tria_mesh.py:1668-1694)."""
        if self.v.shape[0] != vfunc.shape[0]:
            raise ValueError("Error: length of vfunc needs to match number of vertices")
        return np.sum((np.array(vfunc) / 3.0)[self.t], axis=1)

    def smooth_laplace(self, vfunc=None, n: int = 1, lambda_: float = 0.5, mat=None) -> np.ndarray:
        """``n`` steps of v <- (1 - lambda) v + lambda M v with the vertex-area weighted, row-stochastic
        1-ring matrix M (lapy/tria_mesh.py:1696-1715, :1747-1792)."""
        vfunc = np.array(self.v if vfunc is None else vfunc)
        if self.v.shape[0] != vfunc.shape[0]:
            raise ValueError("Error: length of vfunc needs to match number of vertices")
        if mat is None:
            adj = self.adj_sym.copy()
            adj.data = np.ones(adj.data.shape)
            mat = adj.multiply(self.vertex_areas()[:, np.newaxis])
            rowsum = np.sum(mat, axis=1)
            rowsum[rowsum == 0] = 1.0
            mat = mat.multiply(1.0 / rowsum)
        for _ in range(n):
            vfunc = (1.0 - lambda_) * vfunc + lambda_ * mat.dot(vfunc)
        return vfunc

    def curvature(self, smoothit: int = 3):
        """Principal curvature directions / values at the vertices after Alliez et al. 2003
        (lapy/tria_mesh.py:1078-1219): dihedral angle x edge tensor summed over the edges of a vertex,
        ``smoothit`` Laplace smoothing steps, 3x3 eigen-decomposition, the direction closest to the
        vertex normal is the normal.  Returns ``u_min, u_max, c_min, c_max, c_mean, c_gauss, normals``."""
        eps = np.finfo(float).eps
        vids, tids = self.edges()
        tn = self.tria_normals()
        n0, n1 = tn[tids[:, 0], :], tn[tids[:, 1], :]
        angle = np.arccos(np.minimum(np.maximum(np.sum(n0 * n1, axis=1), -1), 1))
        evec = self.v[vids[:, 1], :] - self.v[vids[:, 0], :]
        elen = np.sqrt(np.sum(evec**2, axis=1))
        angle = angle * -np.sign(np.sum(np.cross(n0, n1) * evec, axis=1))  # convex / concave across the edge
        elen[elen < eps] = 1
        evec = evec / elen.reshape(-1, 1)
        elen = elen / np.mean(elen)
        ee = np.empty([elen.shape[0], 6])
        for col, (a, b) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
            ee[:, col] = evec[:, a] * evec[:, b]
        ee = ee * (angle * elen).reshape(-1, 1)
        vnum = self.v.shape[0]
        vv = np.zeros([vnum, 6])
        np.add.at(vv, vids[:, 0], ee)
        np.add.at(vv, vids[:, 1], ee)
        vdeg = np.zeros([vnum])
        np.add.at(vdeg, vids[:, 0], 1)
        np.add.at(vdeg, vids[:, 1], 1)
        vdeg[vdeg == 0] = 1
        vv = self.smooth_laplace(vfunc=vv / vdeg.reshape(-1, 1), n=smoothit, lambda_=1.0)
        mats = np.empty([vnum, 3, 3])
        mats[:, 0, :] = vv[:, [0, 1, 2]]
        mats[:, [1, 2], 0] = vv[:, [1, 2]]
        mats[:, 1, [1, 2]] = vv[:, [3, 4]]
        mats[:, 2, 1] = vv[:, 4]
        mats[:, 2, 2] = vv[:, 5]
        evals, evecs = np.linalg.eigh(mats)
        vnormals = self.vertex_normals()
        along = -np.abs(np.squeeze(np.sum(evecs * vnormals[:, :, np.newaxis], axis=1)))
        i = np.argsort(along, axis=1)
        evals = np.take_along_axis(evals, i, axis=1)
        evecs = np.take_along_axis(evecs, np.tile(i.reshape((vnum, 1, 3)), (1, 3, 1)), axis=2)
        u_min, u_max = np.squeeze(evecs[:, :, 2]), np.squeeze(evecs[:, :, 1])
        c_min, c_max = evals[:, 1], evals[:, 2]
        normals = np.squeeze(evecs[:, :, 0])
        c_mean, c_gauss = (c_min + c_max) / 2.0, c_min * c_max
        swap = np.squeeze(np.where(c_min > c_max))
        c_min[swap], c_max[swap] = c_max[swap], c_min[swap]
        u_min[swap, :], u_max[swap, :] = u_max[swap, :], u_min[swap, :]
        normals = normals * np.sign(np.sum(normals * vnormals, axis=1)).reshape(-1, 1)
        flip = np.squeeze(np.where(np.sum(np.multiply(np.cross(u_min, u_max), normals), axis=1) < 0))
        u_max[flip, :] = -u_max[flip, :]
        return u_min, u_max, c_min, c_max, c_mean, c_gauss, normals

    def curvature_tria(self, smoothit: int = 3):
        """Minimal / maximal curvature directions and values on the triangles (lapy/tria_mesh.py:1221-1279):
        vertex values averaged per triangle, ``u_min`` projected into the triangle plane, ``u_max`` =
        normal x u_min.  This is the input of the anisotropic operator (lapy/solver.py:73-91)."""
        u_min, _u_max, c_min, c_max, _mean, _gauss, _nrm = self.curvature(smoothit)
        tumin = self.map_vfunc_to_tfunc(u_min)
        tcmin, tcmax = self.map_vfunc_to_tfunc(c_min), self.map_vfunc_to_tfunc(c_max)
        tn = self.tria_normals()
        tumin2 = tumin - tn * (np.sum(tn * tumin, axis=1)).reshape(-1, 1)
        tuminl = np.sqrt(np.sum(tumin2 * tumin2, axis=1)).reshape(-1, 1)
        tumin2 = tumin2 / np.maximum(tuminl, 1e-8)
        return tumin2, np.cross(tn, tumin2), tcmin, tcmax

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

    @property
    def adj_sym(self):
        """Symmetric adjacency (edge graph), entry = number of tets sharing the edge
        (lapy/tet_mesh.py:103-122); built on first use."""
        if getattr(self, "_adj_sym_t", None) is not self.t:
            from scipy import sparse

            t = self.t.astype(np.int64)
            pairs = ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))
            i = np.column_stack([t[:, a] for a, b in pairs] + [t[:, b] for a, b in pairs]).reshape(-1)
            j = np.column_stack([t[:, b] for a, b in pairs] + [t[:, a] for a, b in pairs]).reshape(-1)
            n = self.v.shape[0]
            self._adj_sym = sparse.csc_matrix((np.ones(i.shape), (i, j)), shape=(n, n))
            self._adj_sym_t = self.t
        return self._adj_sym


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
