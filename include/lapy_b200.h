/*
 * lapy_b200.h - C ABI of the B200-native backend for LaPy's FEM hot path.
 *
 * LaPy (Deep-MI/LaPy, pure Python) has no FFI of its own; the seam this library plugs into is
 * the Python surface of lapy/solver.py, lapy/heat.py, lapy/diffgeo.py and lapy/shapedna.py
 * (SURVEY.md §8b).  Each entry point below names the reference code it replaces (file:line,
 * relative to the reference root).  INTEGRATION.md shows the ctypes stub a LaPy maintainer
 * would add; lapy_b200/_lib.py is that stub in this repo.
 *
 * Conventions
 *   - every function returns an int status (LB_OK == 0) and records a thread-local message
 *     retrievable with lb_last_error();
 *   - all array arguments are HOST pointers to C-contiguous memory, borrowed for the duration
 *     of the call; outputs are written into caller-allocated arrays;
 *   - device objects (lb_mesh, lb_mat, lb_amg) are opaque handles owned by the library,
 *     released with the matching *_free; a context owns one CUDA stream and a stream-ordered
 *     memory pool, calls on one context are serialised by the caller;
 *   - there is no CPU fallback: lb_ctx_create fails when no CUDA device is usable.
 *   - matrices are symmetric sparse matrices stored as canonical CSC == CSR: fp64 values,
 *     int32 sorted unique indices, explicit zeros kept, exactly what
 *     scipy.sparse.csc_matrix((data,(i,j))) yields for the reference.
 */
#ifndef LAPY_B200_H
#define LAPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lb_ctx lb_ctx;
typedef struct lb_mesh lb_mesh;
typedef struct lb_mat lb_mat;

enum lb_status {
    LB_OK = 0,
    LB_ERR_ARG = 1,         /* -> ValueError in the Python shim            */
    LB_ERR_CUDA = 2,        /* -> RuntimeError                             */
    LB_ERR_NOCONV = 3,      /* -> ArpackNoConvergence-like / RuntimeError  */
    LB_ERR_OOM = 4,         /* -> MemoryError                              */
    LB_ERR_UNSUPPORTED = 5  /* -> NotImplementedError                      */
};

enum lb_dtype { LB_F32 = 0, LB_F64 = 1 };

/* which operator an assembly call builds */
enum lb_fem_kind {
    LB_FEM_TRIA = 0,      /* Solver._fem_tria        lapy/solver.py:105-194 */
    LB_FEM_TRIA_ANISO = 1,/* Solver._fem_tria_aniso  lapy/solver.py:196-308 */
    LB_FEM_TRIA_MASS = 2, /* Solver.fem_tria_mass    lapy/solver.py:310-377 (B only) */
    LB_FEM_TETRA = 3      /* Solver._fem_tetra       lapy/solver.py:379-533 */
};

/* iteration report filled by the solvers */
typedef struct lb_info {
    int32_t iterations;     /* outer iterations performed                         */
    int32_t converged;      /* number of converged columns / eigenpairs           */
    int32_t amg_levels;     /* levels of the preconditioner hierarchy (0 = Jacobi) */
    int32_t reserved;
    double residual;        /* max relative residual at exit                       */
    double setup_ms;        /* device time of the preconditioner setup            */
    double solve_ms;        /* device time of the iteration                       */
} lb_info;

/* ---- context ------------------------------------------------------------------------- */
/* stream: a cudaStream_t to run on (e.g. torch.cuda.current_stream().cuda_stream) or NULL
 * to let the context create its own non-blocking stream. */
int lb_ctx_create(int device, void *stream, lb_ctx **out);
int lb_ctx_destroy(lb_ctx *ctx);
int lb_ctx_sync(lb_ctx *ctx);
const char *lb_last_error(void);

/* Page-locked host memory for large results (the eigenvector array of lb_eigs, the solution block of
 * lb_solve): a page-locked destination is filled by one DMA, a pageable one through the library's
 * staging buffers plus a host copy (lapy/solver.py:713 returns a fresh NumPy array per call; the
 * Python binding keeps a small pool of these blocks behind NumPy arrays).  No context needed. */
int lb_host_alloc(size_t bytes, void **out);
int lb_host_free(void *p);

/* lb_eigs keeps its work blocks ([X P W], A., B.: 24 GB for 2.6M vertices and k = 50) in the context between
 * calls; this returns them to the device's memory pool.  The next lb_eigs allocates again. */
int lb_ctx_release_workspace(lb_ctx *ctx);
const char *lb_version(void);
/* device-time stopwatch on the context's stream (CUDA events) */
int lb_timer_start(lb_ctx *ctx);
int lb_timer_stop(lb_ctx *ctx, double *ms);
/* per-kernel-class device timing (CUDA events around the hot launches): classes are
 * 0 SpMM, 1 Gram (X^T Y), 2 block update (X C), 3 small dense (syevd, coarse solves), 4 column
 * dots, 5 elementwise block kernels.
 * enable clears the records (on: 0 = off, 1 = all classes, 2 = the SpMM class only - two event
 * records per launch are not free, so a timed region records only the class its roofline is quoted
 * on); report fills arrays of 6: launches, device ms, work (algorithmic bytes for classes 0/4/5,
 * flops for 1-3). */
int lb_profile_enable(lb_ctx *ctx, int on);
int lb_profile_report(lb_ctx *ctx, int64_t *count, double *ms, double *work);
/* the records of one class aggregated by launch shape (SpMM: shape0 = columns, shape1 = nnz of the
 * matrix; Gram / update: p, q; small dense: order), largest device time first; arrays of `cap` */
int lb_profile_shapes(lb_ctx *ctx, int cls, int cap, int64_t *shape0, int64_t *shape1, int64_t *count,
                      double *ms, double *work, int *nshapes);
/* number of kernels this library has launched on ctx since creation */
int lb_launch_count(lb_ctx *ctx, int64_t *count);
/* out4: [0] kernels launched, [1] assemblies done by the strip-cooperative triangle kernels, [2] by
 * the record pipeline (tets; triangle meshes with valence > 8, repeated vertices or degenerate
 * elements), [3] reserved.  Lets a test assert that the fast path is the one that runs. */
int lb_ctx_counters(lb_ctx *ctx, int64_t *out4);

/* ---- row-partitioned multi-GPU mode (one process per GPU, NCCL) -------------------------------
 * rank 0 creates an id (128 bytes), the host program broadcasts it (torch.distributed, MPI, ...),
 * every rank calls lb_comm_init on its context.  Afterwards lb_eigs on that context runs
 * row-partitioned over all ranks: every rank passes the same (full) matrices and receives the full
 * result.  The reference has no distributed mode. */
int lb_nccl_unique_id(unsigned char *out128);
int lb_comm_init(lb_ctx *ctx, int world, int rank, const unsigned char *id128);
int lb_comm_destroy(lb_ctx *ctx);
/* self-check of the communicating primitives (row-partitioned SpMM, all-reduced Gram) against
 * redundant full computations on this rank: errs[0], errs[1] = max abs differences */
int lb_dist_selftest(lb_ctx *ctx, lb_mat *a, double *errs);

/* ---- mesh upload: geometry.v / geometry.t as the reference's Solver reads them --------- */
/* v: (nv,3) LB_F32|LB_F64; t: (nt,k) signed integers of t_itemsize 4|8 bytes, k = 3|4.
 * Replaces the fancy-index gathers at lapy/solver.py:145-150, :418-425. */
int lb_mesh_create(lb_ctx *ctx, const void *v, int v_dtype, int64_t nv, const void *t,
                   int t_itemsize, int64_t nt, int k, lb_mesh **out);
int lb_mesh_update_vertices(lb_mesh *mesh, const void *v, int v_dtype);
/* forget the cached vertex->element incidence (rebuilt by the next assembly; benchmarking aid) */
int lb_mesh_drop_cache(lb_mesh *mesh);
int lb_mesh_free(lb_mesh *mesh);

/* ---- assembly -------------------------------------------------------- */
/* kind: lb_fem_kind.  lump != 0 -> diagonal mass.  u1,u2 (nt,3) and aniso_mat (nt,2) fp64 host
 * arrays, only for LB_FEM_TRIA_ANISO.  a_out may be NULL (and is ignored for TRIA_MASS). */
int lb_fem_assemble(lb_ctx *ctx, lb_mesh *mesh, int kind, int lump, const double *u1,
                    const double *u2, const double *aniso_mat, lb_mat **a_out, lb_mat **b_out);

/* ---- matrices --------------------------------------------------------------------------- */
int lb_mat_info(lb_mat *m, int64_t *n, int64_t *nnz);
int lb_mat_download(lb_mat *m, int32_t *indptr, int32_t *indices, double *data);
/* user supplied matrix (e.g. ``fem.mass = sparse.eye(n)``, lapy/diffgeo.py:149) */
int lb_mat_upload(lb_ctx *ctx, int64_t n, int64_t nnz, const int32_t *indptr,
                  const int32_t *indices, const double *data, lb_mat **out);
int lb_mat_free(lb_mat *m);
/* y (n,m) row-major = M x (n,m) row-major; replaces csc_matvec(s) (lapy/solver.py:844-846) */
int lb_spmm(lb_ctx *ctx, lb_mat *mat, const double *x, int64_t m, double *y);

/* device-resident timing of the SpMM kernel (x, y stay in HBM): ms per launch, CUDA events;
 * renumber != 0: in the locality numbering the solvers iterate in (how assembled matrices are
 * stored); 0: the same operator converted to the caller's numbering (what the numbering buys) */
int lb_spmm_benchmark(lb_ctx *ctx, lb_mat *mat, int64_t m, int reps, int renumber,
                      double *ms_per_launch);

/* kernel-level parity of the SpMM forms on resident random blocks of m (% 4 == 0) columns: errs[0..4] = max
 * |strip-staged - row-wise| per epilogue mode in double (same summation order: 0.0 expected), errs[5..9] =
 * max |single-precision strip - double| / max |double| (the multigrid cycle of the preconditioner).
 * The row-wise kernel is the one pinned against scipy's csr_matvecs (lapy/solver.py:844-846) since round 1. */
int lb_spmm_selftest(lb_ctx *ctx, lb_mat *mat, int64_t m, double *errs);

/* dense tall-skinny block products on the fp64 tensor cores (hand-written DMMA kernels), the
 * contractions LAPACK performs inside ARPACK for the reference (lapy/solver.py:713); row-major:
 * C(p,q) = X(n,p)^T Y(n,q)   and   Y(n,q) = alpha X(n,p) C(p,q) + beta Y */
int lb_block_gram(lb_ctx *ctx, int64_t n, int64_t p, const double *x, int64_t q, const double *y,
                  double *cmat);
int lb_block_update(lb_ctx *ctx, int64_t n, int64_t p, const double *x, int64_t q,
                    const double *cmat, double alpha, double beta, double *y);

/* device-resident timing of those kernels (operands stay in HBM): op 0 = Gram, 1 = update -> ms per
 * launch; op 2 = register-only DMMA probe -> sustained fp64 tensor TFLOP/s (the roofline the dense
 * products are quoted against).  variant 1 (A/B aid): the update with its 64-column tile for every q. */
int lb_dense_benchmark(lb_ctx *ctx, int64_t n, int64_t p, int64_t q, int op, int variant, int reps,
                       double *result);

/* ---- solvers ------------------------------------------------------------------------------ */
/* Solver.eigs (lapy/solver.py:667-716): k eigenpairs of A x = lambda B x nearest sigma (<= 0),
 * ascending, B-orthonormal.  evals (k), evecs (n,k) row-major or NULL (eigenvalues only: the
 * ShapeDNA of a batch needs no eigenvectors, lapy/shapedna.py:160-164 keeps them optional downstream).
 * tol <= 0 / maxit <= 0 pick defaults (1e-9 relative residual, 200). */
int lb_eigs(lb_ctx *ctx, lb_mat *a, lb_mat *b, int k, double sigma, double tol, int maxit,
            double *evals, double *evecs, lb_info *info);

/* Solve (alpha*A + beta*B) x = rhs for m right-hand sides with optional Dirichlet rows:
 *   heat.diffusion    lapy/heat.py:208-227      alpha = t, beta = 1, B = lumped mass
 *   Solver.poisson    lapy/solver.py:848-883    alpha = 1, beta = 0, fix = dtup
 * rhs, x: (n,m) row-major.  fix_idx (nfix) rows are eliminated and x there set to
 * fix_val (nfix) (broadcast over columns), the rhs is corrected with -A d like solver.py:846.
 * project_nullspace != 0: the operator is singular with constant null space (closed mesh
 * Poisson, lapy/diffgeo.py:156): rhs and iterates are kept orthogonal to constants. */
int lb_solve(lb_ctx *ctx, lb_mat *a, double alpha, lb_mat *b, double beta, const double *rhs,
             int64_t m, const int64_t *fix_idx, int64_t nfix, const double *fix_val, double tol,
             int maxit, int project_nullspace, double *x, lb_info *info);

/* ---- differential operators ---------------------------------------- */
/* f (nv,nf) row-major -> g (nt,nf,3): tria_compute_gradient lapy/diffgeo.py:222-300,
 * tet_compute_gradient :846-922 */
int lb_gradient(lb_ctx *ctx, lb_mesh *mesh, const double *f, int64_t nf, double *g);
/* x (nt,nf,3) -> d (nv,nf): tria_compute_divergence lapy/diffgeo.py:303-387,
 * tet_compute_divergence :925-1006 */
int lb_divergence(lb_ctx *ctx, lb_mesh *mesh, const double *x, int64_t nf, double *d);
/* flux form of the same quantity on triangle meshes: tria_compute_divergence2 lapy/diffgeo.py:390-469 */
int lb_divergence2(lb_ctx *ctx, lb_mesh *mesh, const double *x, int64_t nf, double *d);
/* f (nv,nf) -> d (nv,nf) = divergence of the normalised gradient of f, all on the device: the
 * right-hand side of compute_geodesic_f lapy/diffgeo.py:144-156 (gradient, g/|g| with nan_to_num,
 * integrated divergence) */
int lb_unit_gradient_divergence(lb_ctx *ctx, lb_mesh *mesh, const double *f, int64_t nf, double *d);
/* mean length of the unique edges of the mesh `pattern` (a stiffness matrix) was assembled on:
 * TriaMesh.avg_edge_length lapy/tria_mesh.py:735-748, TetMesh lapy/tet_mesh.py:182-195 (fp64) */
int lb_avg_edge_length(lb_ctx *ctx, lb_mesh *mesh, lb_mat *pattern, double *out);
#ifdef __cplusplus
}
#endif
#endif /* LAPY_B200_H */
