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
    """Upload (and cache on the geometry object) v / t."""
    cached = getattr(geometry, "_lb_device_mesh", None)
    if (
        cached is not None
        and cached[0] is geometry.v
        and cached[1] is geometry.t
        and cached[2].ctx is ctx
        and cached[2].handle
    ):
        return cached[2]
    dm = _lib.DeviceMesh(ctx, geometry.v, geometry.t)
    try:
        geometry._lb_device_mesh = (geometry.v, geometry.t, dm)
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
        self._mesh = _device_mesh(geometry, self._ctx)
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
