"""Drop-in for ``lapy.Solver`` (reference lapy/solver.py) backed by liblapyb200.so.

Same constructor, attributes, methods, defaults, return types and exception types as the
reference class; the arithmetic runs on one B200 through the C ABI in
include/lapy_b200.h.  ``.stiffness`` / ``.mass`` are real ``scipy.sparse.csc_matrix`` objects
(fp64, int32 sorted canonical indices, explicit zeros kept) downloaded lazily from the device
CSR, and they stay assignable (lapy/diffgeo.py:149 overwrites ``.mass``): assigning uploads the
new matrix on next use.
"""

from __future__ import annotations

import logging

import numpy as np
from scipy import sparse

from . import _lib

logger = logging.getLogger(__name__)

_KIND = {"TriaMesh": _lib.FEM_TRIA, "TetMesh": _lib.FEM_TETRA}


def _device_mesh(geometry, ctx):
    """Upload ``geometry.v`` / ``geometry.t``.

    The reference always reads the current host arrays, and an in-place edit (``mesh.v *= s``) does
    not change array identity, so a device copy is only re-used when BOTH arrays are read-only
    (``arr.flags.writeable = False``: callers that solve many problems on one mesh opt in that way);
    otherwise every call uploads (19 ms at 2.6M vertices - cheaper than hashing the arrays)."""
    v, t = geometry.v, geometry.t
    frozen = not getattr(getattr(v, "flags", None), "writeable", True) and not getattr(
        getattr(t, "flags", None), "writeable", True
    )
    if frozen:
        cached = getattr(geometry, "_lb_device_mesh", None)
        if cached is not None and cached[0] is v and cached[1] is t and cached[2].ctx is ctx and cached[2].handle:
            return cached[2]
    dm = _lib.DeviceMesh(ctx, v, t)
    if frozen:
        try:
            geometry._lb_device_mesh = (v, t, dm)
        except AttributeError:  # geometry with __slots__
            pass
    return dm


class Solver:
    """FEM solver for the Laplace(-Beltrami) eigenproblem and Poisson equation on a B200.

    Parameters are those of the reference (lapy/solver.py:54-62).  ``use_cholmod`` is accepted
    and ignored (there is no factorisation on the device path); ``dtype`` other than float64
    casts the assembled fp64 matrices.  ``device`` / ``ctx`` select the GPU (extension).
    """

    def __init__(
        self,
        geometry,
        lump: bool = False,
        aniso=None,
        aniso_smooth: int = 10,
        use_cholmod: bool = False,
        dtype=np.float64,
        *,
        device: int | None = None,
        ctx: _lib.Context | None = None,
        _mesh: _lib.DeviceMesh | None = None,
    ) -> None:
        self.sksparse = None
        self._dtype = np.dtype(dtype)
        if self._dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("dtype must be float32 or float64")
        self._ctx = ctx if ctx is not None else _lib.default_context(device)
        name = type(geometry).__name__
        if name == "TriaMesh":
            if aniso is not None:
                logger.info("TriaMesh with anisotropic Laplace-Beltrami")
                if not hasattr(geometry, "curvature_tria"):
                    raise NotImplementedError(
                        "aniso needs geometry.curvature_tria() (lapy.TriaMesh provides it; SURVEY.md §8f)"
                    )
                u1, u2, c1, c2 = geometry.curvature_tria(smoothit=aniso_smooth)
                if isinstance(aniso, (list, tuple, set, np.ndarray)):
                    if len(aniso) != 2:
                        raise ValueError("aniso should be scalar or tuple/array of length 2!")
                    aniso0, aniso1 = aniso[0], aniso[1]
                else:
                    aniso0 = aniso1 = aniso
                aniso_mat = np.empty((geometry.t.shape[0], 2), dtype=self._dtype)
                aniso_mat[:, 1] = np.exp(-aniso1 * np.abs(c1))
                aniso_mat[:, 0] = np.exp(-aniso0 * np.abs(c2))
                kind, extra = _lib.FEM_TRIA_ANISO, (u1, u2, aniso_mat)
            else:
                logger.info("TriaMesh with regular Laplace-Beltrami")
                kind, extra = _lib.FEM_TRIA, None
        elif name == "TetMesh":
            logger.info("TetMesh with regular Laplace")
            kind, extra = _lib.FEM_TETRA, None
        else:
            raise ValueError('Geometry type "' + name + '" unknown')
        self._mesh = _mesh if _mesh is not None and _mesh.ctx is self._ctx else _device_mesh(geometry, self._ctx)
        self._dev = dict(zip("ab", _lib.assemble(self._ctx, self._mesh, kind, lump, extra)))
        self._host = {"a": None, "b": None}
        self.geotype = type(geometry)
        self.use_cholmod = use_cholmod
        self.lump = bool(lump)

    # -- matrices: lazy device -> SciPy, assignable ------------------------------------------
    def _get(self, which):
        if self._host[which] is None:
            m = self._dev[which].to_scipy()
            if self._dtype != np.float64:
                m = m.astype(self._dtype)
            self._host[which] = m
        return self._host[which]

    def _set(self, which, value):
        self._host[which] = value
        self._dev[which] = None  # stale: re-uploaded on next device use

    def _device(self, which) -> _lib.DeviceMatrix:
        if self._dev[which] is None:
            self._dev[which] = _lib.DeviceMatrix.from_scipy(self._ctx, self._host[which])
        return self._dev[which]

    stiffness = property(lambda self: self._get("a"), lambda self, m: self._set("a", m))
    mass = property(lambda self: self._get("b"), lambda self, m: self._set("b", m))

    # -- static assembly entry points (lapy/solver.py:105, :196, :310, :379) -------------------
    @staticmethod
    def _assemble_static(geometry, kind, lump, dtype, extra=None, want_a=True):
        ctx = _lib.default_context()
        mesh = _device_mesh(geometry, ctx)
        a, b = _lib.assemble(ctx, mesh, kind, lump, extra, want_a=want_a)
        dtype = np.dtype(dtype)
        out = [m.to_scipy() if m is not None else None for m in (a, b)]
        if dtype != np.float64:
            out = [m.astype(dtype) if m is not None else None for m in out]
        return out

    @staticmethod
    def _fem_tria(tria, lump: bool = False, dtype=np.float64):
        return tuple(Solver._assemble_static(tria, _lib.FEM_TRIA, lump, dtype))

    @staticmethod
    def _fem_tria_aniso(tria, u1, u2, aniso_mat, lump: bool = False, dtype=np.float64):
        return tuple(Solver._assemble_static(tria, _lib.FEM_TRIA_ANISO, lump, dtype, (u1, u2, aniso_mat)))

    @staticmethod
    def fem_tria_mass(tria, lump: bool = False, dtype=np.float64):
        return Solver._assemble_static(tria, _lib.FEM_TRIA_MASS, lump, dtype, want_a=False)[1]

    @staticmethod
    def _fem_tetra(tetra, lump: bool = False, dtype=np.float64):
        return tuple(Solver._assemble_static(tetra, _lib.FEM_TETRA, lump, dtype))

    # -- Solver.eigs (lapy/solver.py:667-716) ----------------------------------------------------
    def eigs(self, k: int = 10, sigma: float = -0.01, *, tol: float = 0.0, maxit: int = 0, vectors: bool = True):
        """k eigenpairs of ``A x = lambda B x`` nearest ``sigma`` (sigma <= 0: the k smallest).

        Returns ``(eigenvalues (k,), eigenvectors (n, k))``, ascending, B-orthonormal like ARPACK's.
        ``tol`` / ``maxit`` (extensions) bound the block-LOBPCG iteration; ``self.last_info`` holds
        the iteration report.  sigma > 0 raises ``NotImplementedError``.  ``vectors=False``
        (extension) skips the eigenvector download and returns ``(eigenvalues, None)``.
        """
        n = self._shape0()
        if k >= n:  # SciPy ARPACK wrapper raises the same for sparse input (arpack.py:1691-1699)
            raise TypeError(f"Cannot use scipy.linalg.eigh for sparse A with k >= N. k={k}, N={n}")
        if k <= 0:
            raise ValueError(f"k must be greater than 0. k={k}")
        logger.info("Solver: block LOBPCG + smoothed-aggregation AMG on the GPU ...")
        evals, evecs, info = _lib.eigs(self._ctx, self._device("a"), self._device("b"), k, sigma, tol, maxit,
                                       vectors=vectors)
        self.last_info = info
        return evals, evecs

    eigensystem = eigs  # name used by README.md:51 / BASELINE.json

    def _shape0(self):
        for which in "ab":
            if self._dev[which] is not None:
                return self._dev[which].n
        return self._host["a"].shape[0]

    # -- Solver.poisson (lapy/solver.py:718-889) -----------------------------------------------------
    def poisson(self, h=0.0, dtup=(), ntup=(), *, tol: float = 0.0, maxit: int = 0):
        """Solve ``A x = B (h - n) - A d`` with Dirichlet (``dtup``) / Neumann (``ntup``) data.

        Same argument checks, broadcasting and return shapes as the reference.  The solve is an
        AMG-preconditioned block CG over all right-hand sides; without Dirichlet data the operator
        is singular (constants) and the zero-mean solution is returned (the reference returns the
        LU solution, defined up to the same constant).
        """
        a_dev = self._device("a")
        dim = a_dev.n
        b_dev = self._device("b")
        if b_dev.n != dim:
            raise ValueError("Error: Square input matrices should have same number of rows and columns.")
        dtype = np.float64
        if np.isscalar(h):
            h = np.full((dim, 1), h, dtype=dtype)
        else:
            h = np.asarray(h, dtype=dtype)
            if h.ndim == 1:
                if h.size != dim:
                    raise ValueError("h should be either scalar or column vector with row num of A")
                h = h[:, np.newaxis]
            elif h.ndim == 2:
                if h.shape[0] != dim:
                    raise ValueError("h should be either scalar or array with first dim matching A")
            else:
                raise ValueError("h should be either scalar or 1-D/2-D array")
        n_rhs = h.shape[1]
        scalar_rhs = n_rhs == 1
        didx, ddat = [], []
        if dtup:
            if len(dtup) != 2:
                raise ValueError("dtup should contain index and data arrays")
            didx, ddat = dtup[0], dtup[1]
            if np.unique(didx).size != len(didx):
                raise ValueError("dtup indices need to be unique")
            if not (len(didx) > 0 and len(didx) == len(ddat)):
                raise ValueError("dtup should contain index and data arrays (same lengths > 0)")
        if ntup:
            if len(ntup) != 2:
                raise ValueError("ntup should contain index and data arrays")
            nidx, ndat = ntup[0], ntup[1]
            if not (len(nidx) > 0 and len(nidx) == len(ndat)):
                raise ValueError("ntup should contain index and data arrays (same lengths > 0)")
            h = h.copy()
            np.subtract.at(h, (np.asarray(nidx), slice(None)), np.asarray(ndat, dtype=dtype)[:, None])
        logger.info("Solver: AMG-preconditioned block CG on the GPU ...")
        rhs = _lib.spmm(self._ctx, b_dev, h)  # b = M (h - n)
        x, info = _lib.solve(
            self._ctx, a_dev, 1.0, None, 0.0, rhs,
            fix_idx=np.asarray(didx) if len(didx) else None,
            fix_val=np.asarray(ddat, dtype=dtype) if len(didx) else None,
            tol=tol, maxit=maxit, project_nullspace=len(didx) == 0,
        )  # fmt: skip
        self.last_info = info
        if scalar_rhs:
            return np.squeeze(np.array(x))
        return np.array(x)
