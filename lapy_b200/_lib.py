"""ctypes binding of liblapyb200.so (include/lapy_b200.h) - the only way host Python reaches
the device code.  There is NO CPU fallback: if the shared library is missing or no CUDA device
is usable, importing / first use fails loudly (``ImportError`` / ``RuntimeError``).
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblapyb200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOCONV, ERR_OOM, ERR_UNSUPPORTED = range(6)
F32, F64 = 0, 1
FEM_TRIA, FEM_TRIA_ANISO, FEM_TRIA_MASS, FEM_TETRA = range(4)


class NoConvergence(RuntimeError):
    """Iterative solver did not reach its tolerance (cf. scipy ArpackNoConvergence)."""


class Info(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32),
        ("converged", C.c_int32),
        ("amg_levels", C.c_int32),
        ("reserved", C.c_int32),
        ("residual", C.c_double),
        ("setup_ms", C.c_double),
        ("solve_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_vp, _i64, _int, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_pp = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int status except the two string getters
SIGNATURES = {
    "lb_ctx_create": [_int, _vp, _pp],
    "lb_ctx_destroy": [_vp],
    "lb_ctx_sync": [_vp],
    "lb_timer_start": [_vp],
    "lb_timer_stop": [_vp, C.POINTER(_dbl)],
    "lb_launch_count": [_vp, C.POINTER(_i64)],
    "lb_ctx_counters": [_vp, _vp],
    "lb_profile_enable": [_vp, _int],
    "lb_profile_report": [_vp, _vp, _vp, _vp],
    "lb_profile_shapes": [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, C.POINTER(_int)],
    "lb_nccl_unique_id": [_vp],
    "lb_comm_init": [_vp, _int, _int, _vp],
    "lb_comm_destroy": [_vp],
    "lb_dist_selftest": [_vp, _vp, _vp],
    "lb_mesh_create": [_vp, _vp, _int, _i64, _vp, _int, _i64, _int, _pp],
    "lb_mesh_update_vertices": [_vp, _vp, _int],
    "lb_mesh_drop_cache": [_vp],
    "lb_mesh_free": [_vp],
    "lb_fem_assemble": [_vp, _vp, _int, _int, _vp, _vp, _vp, _pp, _pp],
    "lb_mat_info": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "lb_mat_download": [_vp, _vp, _vp, _vp],
    "lb_mat_upload": [_vp, _i64, _i64, _vp, _vp, _vp, _pp],
    "lb_mat_free": [_vp],
    "lb_spmm": [_vp, _vp, _vp, _i64, _vp],
    "lb_spmm_benchmark": [_vp, _vp, _i64, _int, _int, C.POINTER(_dbl)],
    "lb_block_gram": [_vp, _i64, _i64, _vp, _i64, _vp, _vp],
    "lb_block_update": [_vp, _i64, _i64, _vp, _i64, _vp, _dbl, _dbl, _vp],
    "lb_dense_benchmark": [_vp, _i64, _i64, _i64, _int, _int, _int, C.POINTER(_dbl)],
    "lb_ctx_release_workspace": [_vp],
    "lb_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "lb_host_free": [_vp],
    "lb_spmm_selftest": [_vp, _vp, _i64, C.POINTER(C.c_double)],
    "lb_eigs": [_vp, _vp, _vp, _int, _dbl, _dbl, _int, _vp, _vp, C.POINTER(Info)],
    "lb_solve": [_vp, _vp, _dbl, _vp, _dbl, _vp, _i64, _vp, _i64, _vp, _dbl, _int, _int, _vp, C.POINTER(Info)],
    "lb_avg_edge_length": [_vp, _vp, _vp, C.POINTER(_dbl)],
    "lb_gradient": [_vp, _vp, _vp, _i64, _vp],
    "lb_divergence": [_vp, _vp, _vp, _i64, _vp],
    "lb_divergence2": [_vp, _vp, _vp, _i64, _vp],
    "lb_unit_gradient_divergence": [_vp, _vp, _vp, _i64, _vp],
}
STRING_GETTERS = ("lb_last_error", "lb_version")

_lib = None
_lock = threading.Lock()


def lib():
    """The loaded shared library (loaded once)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (nvcc, sm_100a). lapy_b200 has no CPU fallback."
                    )
                handle = C.CDLL(LIB_PATH)
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.argtypes = argtypes
                    fn.restype = C.c_int
                for name in STRING_GETTERS:
                    getattr(handle, name).restype = C.c_char_p
                    getattr(handle, name).argtypes = []
                _lib = handle
    return _lib


def check(status: int):
    if status == OK:
        return
    msg = lib().lb_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(msg)
    if status == ERR_NOCONV:
        raise NoConvergence(msg)
    if status == ERR_OOM:
        raise MemoryError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """A page-locked host block (lb_host_alloc) that NumPy can wrap (``__array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class _PinnedPool:
    """Large result arrays (eigenvectors: 1 GB at level 9) are NumPy arrays on page-locked blocks: the
    library fills them with one DMA instead of staging + host copy, and a loop that drops its previous
    result gets the same block back (no gigabyte of page faults per call).  At most ``max_bytes`` are
    held (live + free); beyond that, and for small arrays, plain ``np.empty``."""

    max_bytes = 4 << 30
    min_bytes = 32 << 20

    def __init__(self):
        import threading

        self.lock = threading.Lock()
        self.free: dict[int, list[int]] = {}
        self.total = 0

    def _release(self, ptr: int, nbytes: int):
        with self.lock:
            self.free.setdefault(nbytes, []).append(ptr)

    def empty(self, shape, dtype=np.float64) -> np.ndarray:
        import weakref

        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if nbytes < self.min_bytes:
            return np.empty(shape, dtype)
        ptr = None
        with self.lock:
            if self.free.get(nbytes):
                ptr = self.free[nbytes].pop()
            elif self.total + nbytes > self.max_bytes:
                for size, blocks in list(self.free.items()):  # make room from free blocks of other sizes
                    while blocks and self.total + nbytes > self.max_bytes:
                        lib().lb_host_free(C.c_void_p(blocks.pop()))
                        self.total -= size
                if self.total + nbytes > self.max_bytes:
                    return np.empty(shape, dtype)
            if ptr is None:
                self.total += nbytes
        if ptr is None:
            out = C.c_void_p()
            if lib().lb_host_alloc(nbytes, C.byref(out)) != 0 or not out.value:
                with self.lock:
                    self.total -= nbytes
                return np.empty(shape, dtype)
            ptr = out.value
        block = _PinnedBlock(ptr, nbytes)
        fin = weakref.finalize(block, self._release, ptr, nbytes)
        fin.atexit = False
        return np.asarray(block).view(dtype).reshape(shape)


_pinned = _PinnedPool()


class Context:
    """One CUDA stream + memory pool on one device (lb_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        check(lib().lb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            lib().lb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().lb_ctx_sync(self.handle))

    def release_workspace(self):
        """Give the eigensolver's persistent work blocks (24 GB after a 2.6M-vertex, k=50 solve) back to the
        device's memory pool; the next ``eigs`` on this context allocates them again."""
        check(lib().lb_ctx_release_workspace(self.handle))

    def init_row_partition(self):
        """Join the row-partitioned mode: all ranks of the current torch.distributed group create
        one NCCL communicator inside the library (lb_comm_init); afterwards ``Solver.eigs`` on this
        context runs one eigensolve cooperatively over all ranks."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            check(lib().lb_nccl_unique_id(ptr(buf)))
        t = torch.from_numpy(buf)
        if dist.get_backend() == "nccl":
            t = t.to(torch.device("cuda", self.device))
        dist.broadcast(t, 0)
        buf = np.ascontiguousarray(t.cpu().numpy())
        check(lib().lb_comm_init(self.handle, world, rank, ptr(buf)))
        self.world, self.rank = world, rank

    def init_row_partition_single(self):
        """One-rank communicator (testing aid: a context with a communicator runs the row-partitioned
        code path, here on one GPU)."""
        buf = np.zeros(128, np.uint8)
        check(lib().lb_nccl_unique_id(ptr(buf)))
        check(lib().lb_comm_init(self.handle, 1, 0, ptr(buf)))
        self.world, self.rank = 1, 0

    def leave_row_partition(self):
        check(lib().lb_comm_destroy(self.handle))

    def timer_start(self):
        check(lib().lb_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(lib().lb_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    PROFILE_CLASSES = ("spmm", "gram", "update", "small_dense", "col_dots", "elementwise")

    def profile_enable(self, on: bool | int = True):
        """0 / False: off; 1 / True: every kernel class; 2: the SpMM class only (cheap enough for a timed region)."""
        check(lib().lb_profile_enable(self.handle, int(on)))

    def profile_report(self) -> dict:
        """{class: {launches, ms, work}} since profile_enable(True); work = bytes or flops."""
        cnt = np.zeros(6, np.int64)
        ms = np.zeros(6, np.float64)
        work = np.zeros(6, np.float64)
        check(lib().lb_profile_report(self.handle, ptr(cnt), ptr(ms), ptr(work)))
        return {
            name: {"launches": int(cnt[i]), "ms": float(ms[i]), "work": float(work[i])}
            for i, name in enumerate(self.PROFILE_CLASSES)
            if cnt[i]
        }

    def profile_shapes(self, cls: str, cap: int = 16) -> list[dict]:
        """Records of one class aggregated by launch shape, largest device time first."""
        s0, s1, cnt = (np.zeros(cap, np.int64) for _ in range(3))
        ms, work = np.zeros(cap), np.zeros(cap)
        n = C.c_int()
        check(lib().lb_profile_shapes(self.handle, self.PROFILE_CLASSES.index(cls), cap, ptr(s0), ptr(s1), ptr(cnt),
                                      ptr(ms), ptr(work), C.byref(n)))  # fmt: skip
        return [
            {"shape": (int(s0[i]), int(s1[i])), "launches": int(cnt[i]), "ms": float(ms[i]), "work": float(work[i])}
            for i in range(n.value)
        ]

    def counters(self) -> dict:
        """{launches, strip_assemblies, record_assemblies} since the context was created."""
        out = np.zeros(4, np.int64)
        check(lib().lb_ctx_counters(self.handle, ptr(out)))
        return {"launches": int(out[0]), "strip_assemblies": int(out[1]), "record_assemblies": int(out[2])}

    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib().lb_launch_count(self.handle, C.byref(n)))
        return n.value


_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    """Process-wide context per device; device defaults to $LAPY_B200_DEVICE, $LOCAL_RANK or 0."""
    if device is None:
        device = int(os.environ.get("LAPY_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceMesh:
    """geometry.v / geometry.t resident on the device (lb_mesh)."""

    def __init__(self, ctx: Context, v: np.ndarray, t: np.ndarray):
        v = np.asarray(v)
        t = np.asarray(t)
        if v.ndim != 2 or v.shape[1] != 3:
            raise ValueError("vertices must have shape (n, 3)")
        if t.ndim != 2 or t.shape[1] not in (3, 4):
            raise ValueError("elements must have shape (m, 3) or (m, 4)")
        if v.dtype != np.float32:
            v = v.astype(np.float64, copy=False)
        if t.dtype.kind not in "iu":
            raise ValueError("element indices must be integers")
        if t.dtype.itemsize not in (4, 8) or t.dtype.kind == "u" or not t.dtype.isnative:
            t = t.astype(np.int64)  # big-endian FreeSurfer '>i4', uint*, int16 ...
        v = np.ascontiguousarray(v)  # TriaMesh may hand over a transposed view
        t = np.ascontiguousarray(t)
        self.ctx = ctx
        self.nv, self.nt, self.k = v.shape[0], t.shape[0], t.shape[1]
        self.v_dtype = v.dtype
        h = C.c_void_p()
        check(
            lib().lb_mesh_create(
                ctx.handle, ptr(v), F32 if v.dtype == np.float32 else F64, self.nv,
                ptr(t), t.dtype.itemsize, self.nt, self.k, C.byref(h),
            )
        )  # fmt: skip
        self.handle = h

    def drop_cache(self):
        check(lib().lb_mesh_drop_cache(self.handle))

    def update_vertices(self, v: np.ndarray):
        """New vertex positions on the same connectivity (lb_mesh_update_vertices): loop callers such
        as the mean curvature flow re-assemble without re-uploading the elements."""
        v = np.asarray(v)
        if v.shape != (self.nv, 3):
            raise ValueError("vertices must keep the shape (n, 3)")
        if v.dtype != np.float32:
            v = v.astype(np.float64, copy=False)
        v = np.ascontiguousarray(v)
        check(lib().lb_mesh_update_vertices(self.handle, ptr(v), F32 if v.dtype == np.float32 else F64))
        self.v_dtype = v.dtype

    def close(self):
        if getattr(self, "handle", None):
            lib().lb_mesh_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMatrix:
    """Symmetric sparse matrix in canonical CSC==CSR on the device (lb_mat)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.handle = handle
        n, nnz = C.c_int64(), C.c_int64()
        check(lib().lb_mat_info(handle, C.byref(n), C.byref(nnz)))
        self.n, self.nnz = n.value, nnz.value

    @classmethod
    def from_scipy(cls, ctx: Context, m):
        from scipy import sparse

        m = sparse.csc_matrix(m)
        if m.shape[0] != m.shape[1]:
            raise ValueError("matrix must be square")
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data, dtype=np.float64)
        h = C.c_void_p()
        check(lib().lb_mat_upload(ctx.handle, m.shape[0], m.nnz, ptr(indptr), ptr(indices), ptr(data), C.byref(h)))
        return cls(ctx, h)

    def to_scipy(self):
        from scipy import sparse

        indptr = np.empty(self.n + 1, np.int32)
        indices = np.empty(self.nnz, np.int32)
        data = np.empty(self.nnz, np.float64)
        check(lib().lb_mat_download(self.handle, ptr(indptr), ptr(indices), ptr(data)))
        m = sparse.csc_matrix((data, indices, indptr), shape=(self.n, self.n))
        m.has_canonical_format = True
        return m

    def close(self):
        if getattr(self, "handle", None):
            lib().lb_mat_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def assemble(ctx: Context, mesh: DeviceMesh, kind: int, lump: bool, aniso=None, want_a: bool = True):
    """lb_fem_assemble -> (DeviceMatrix A | None, DeviceMatrix B)."""
    u1 = u2 = am = None
    if kind == FEM_TRIA_ANISO:
        u1, u2, am = (np.ascontiguousarray(x, dtype=np.float64) for x in aniso)
        if u1.shape != (mesh.nt, 3) or u2.shape != (mesh.nt, 3) or am.shape != (mesh.nt, 2):
            raise ValueError("u1, u2 must have shape (n_triangles, 3) and aniso_mat (n_triangles, 2)")
    ha, hb = C.c_void_p(), C.c_void_p()
    check(
        lib().lb_fem_assemble(
            ctx.handle, mesh.handle, kind, int(bool(lump)), ptr(u1), ptr(u2), ptr(am),
            C.byref(ha) if want_a else None, C.byref(hb),
        )
    )  # fmt: skip
    a = DeviceMatrix(ctx, ha) if (want_a and ha.value) else None
    return a, DeviceMatrix(ctx, hb)


def spmm(ctx: Context, mat: DeviceMatrix, x: np.ndarray) -> np.ndarray:
    """y = M x for x (n,) or (n, m) (lb_spmm)."""
    x = np.asarray(x, dtype=np.float64)
    one_d = x.ndim == 1
    x2 = np.ascontiguousarray(x.reshape(mat.n, -1))
    y = np.empty_like(x2)
    check(lib().lb_spmm(ctx.handle, mat.handle, ptr(x2), x2.shape[1], ptr(y)))
    return y[:, 0] if one_d else y


def eigs(ctx: Context, a: DeviceMatrix, b: DeviceMatrix, k: int, sigma: float, tol: float = 0.0, maxit: int = 0,
         out_evecs: np.ndarray | None = None, vectors: bool = True):
    """lb_eigs -> (evals (k,), evecs (n,k), info dict).  ``out_evecs``: a C-contiguous float64 (n, k) array
    to receive the eigenvectors (a throughput loop re-uses one buffer instead of allocating - and page
    faulting - a fresh gigabyte per call).  ``vectors=False``: eigenvalues only (evecs is None)."""
    evals = np.empty(k, np.float64)
    if not vectors:
        info = Info()
        check(lib().lb_eigs(ctx.handle, a.handle, b.handle, int(k), float(sigma), float(tol), int(maxit),
                            ptr(evals), None, C.byref(info)))  # fmt: skip
        return evals, None, info.as_dict()
    if out_evecs is None:
        evecs = _pinned.empty((a.n, k), np.float64)
    else:
        evecs = out_evecs
        if evecs.shape != (a.n, k) or evecs.dtype != np.float64 or not evecs.flags.c_contiguous:
            raise ValueError("out_evecs must be a C-contiguous float64 array of shape (n, k)")
    info = Info()
    check(lib().lb_eigs(ctx.handle, a.handle, b.handle, int(k), float(sigma), float(tol), int(maxit),
                        ptr(evals), ptr(evecs), C.byref(info)))  # fmt: skip
    return evals, evecs, info.as_dict()


def solve(ctx: Context, a: DeviceMatrix, alpha: float, b: DeviceMatrix | None, beta: float, rhs: np.ndarray,
          fix_idx=None, fix_val=None, tol: float = 0.0, maxit: int = 0, project_nullspace: bool = False):
    """lb_solve: (alpha*A + beta*B) x = rhs, rhs (n, m) -> (x (n, m), info dict)."""
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    if rhs.ndim != 2 or rhs.shape[0] != a.n:
        raise ValueError("rhs must have shape (n, m)")
    x = np.empty_like(rhs)
    nfix = 0
    if fix_idx is not None and len(fix_idx):
        fix_idx = np.ascontiguousarray(fix_idx, dtype=np.int64)
        fix_val = np.ascontiguousarray(fix_val, dtype=np.float64)
        nfix = len(fix_idx)
    else:
        fix_idx = fix_val = None
    info = Info()
    check(lib().lb_solve(ctx.handle, a.handle, float(alpha), b.handle if b is not None else None, float(beta),
                         ptr(rhs), rhs.shape[1], ptr(fix_idx), nfix, ptr(fix_val), float(tol), int(maxit),
                         int(bool(project_nullspace)), ptr(x), C.byref(info)))  # fmt: skip
    return x, info.as_dict()


def gradient(ctx: Context, mesh: DeviceMesh, f: np.ndarray) -> np.ndarray:
    """f (nv, nf) -> (nt, nf, 3) (lb_gradient)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    g = np.empty((mesh.nt, f.shape[1], 3), np.float64)
    check(lib().lb_gradient(ctx.handle, mesh.handle, ptr(f), f.shape[1], ptr(g)))
    return g


def divergence(ctx: Context, mesh: DeviceMesh, x: np.ndarray, flux: bool = False) -> np.ndarray:
    """x (nt, nf, 3) -> (nv, nf) (lb_divergence; flux=True: lb_divergence2, triangles only)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    d = np.empty((mesh.nv, x.shape[1]), np.float64)
    fn = lib().lb_divergence2 if flux else lib().lb_divergence
    check(fn(ctx.handle, mesh.handle, ptr(x), x.shape[1], ptr(d)))
    return d


def unit_gradient_divergence(ctx: Context, mesh: DeviceMesh, f: np.ndarray) -> np.ndarray:
    """f (nv, nf) -> div(grad f / |grad f|) (nv, nf) on the device (lb_unit_gradient_divergence)."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    d = np.empty((mesh.nv, f.shape[1]), np.float64)
    check(lib().lb_unit_gradient_divergence(ctx.handle, mesh.handle, ptr(f), f.shape[1], ptr(d)))
    return d


def block_gram(ctx: Context, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """x^T y for tall-skinny row-major blocks (lb_block_gram)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((x.shape[1], y.shape[1]), np.float64)
    check(lib().lb_block_gram(ctx.handle, x.shape[0], x.shape[1], ptr(x), y.shape[1], ptr(y), ptr(out)))
    return out


def block_update(ctx: Context, x: np.ndarray, cmat: np.ndarray, alpha=1.0, beta=0.0, y=None) -> np.ndarray:
    """alpha * x @ cmat + beta * y (lb_block_update)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    cmat = np.ascontiguousarray(cmat, dtype=np.float64)
    out = np.zeros((x.shape[0], cmat.shape[1])) if y is None else np.array(y, dtype=np.float64, order="C")
    check(lib().lb_block_update(ctx.handle, x.shape[0], x.shape[1], ptr(x), cmat.shape[1], ptr(cmat),
                                float(alpha), float(beta), ptr(out)))  # fmt: skip
    return out


def spmm_benchmark(ctx: Context, mat: DeviceMatrix, m: int = 1, reps: int = 20, renumber: bool = False,
                   variant: int = 0) -> float:
    """Device time (ms) of one y = M x launch with m columns, x / y resident in HBM; renumber: in the
    solver (locality) numbering instead of the caller's.  variant (development aid): 1 = the row-wise
    fallback kernel, 2 = the single-precision strip kernel of the preconditioner."""
    ms = C.c_double()
    check(lib().lb_spmm_benchmark(ctx.handle, mat.handle, int(m), int(reps), int(bool(renumber)) | (int(variant) << 8),
                                  C.byref(ms)))
    return ms.value


def spmm_selftest(ctx: Context, mat: DeviceMatrix, m: int = 64) -> np.ndarray:
    """(10,) errors of lb_spmm_selftest: [0:5] strip-staged vs row-wise SpMM per epilogue mode (double, expected
    0.0), [5:10] single-precision strip kernel vs double, relative."""
    errs = np.zeros(10)
    check(lib().lb_spmm_selftest(ctx.handle, mat.handle, int(m), errs.ctypes.data_as(C.POINTER(C.c_double))))
    return errs


def dense_benchmark(ctx: Context, n: int, p: int, q: int, op: int, variant: int = 0, reps: int = 10) -> float:
    """ms per launch of the Gram (op 0) / update (op 1) kernel on resident operands; op 2: sustained
    fp64 tensor TFLOP/s of a register-only DMMA probe (lb_dense_benchmark)."""
    out = C.c_double()
    check(lib().lb_dense_benchmark(ctx.handle, int(n), int(p), int(q), int(op), int(variant), int(reps), C.byref(out)))
    return out.value


def avg_edge_length(ctx: Context, mesh: DeviceMesh, pattern: DeviceMatrix) -> float:
    """Mean unique-edge length on the device (fp64 meshes; lb_avg_edge_length)."""
    out = C.c_double()
    check(lib().lb_avg_edge_length(ctx.handle, mesh.handle, pattern.handle, C.byref(out)))
    return out.value
