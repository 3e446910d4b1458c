// blockvec.cuh - sparse x block-vector products, block BLAS-1 and the dense tall-skinny products.
//
// Block vectors are row-major (n, cols) fp64 arrays with a leading dimension ld >= cols: the
// layout the reference returns eigenvectors in (lapy/solver.py:713, C-order (n, k)) and the one
// that makes CSR x block products coalesced (a row of X is one contiguous 8*cols-byte segment).
#pragma once
#include "common.cuh"

namespace lb {

// Y(n,m) = A X                       (mode 0)
// Y(n,m) = B - A X                   (mode 1, residual; B may alias Y)
// Y(n,m) = B + A X                   (mode 2; B may alias Y)
// Y(n,m) = B - A X and D = c2 * dinv o Y               (mode 3, fused first Chebyshev step)
// SOL (+)= X + c1 X + c2 * dinv o (B - A X)              (mode 4, fused last Chebyshev step; Y unused)
template <typename T>
struct SpmmEpilogueT {
    const T *dinv = nullptr;
    T *out2 = nullptr;  // D (mode 3) / SOL (mode 4)
    int ldout2 = 0;
    T c1 = 0, c2 = 0;
    int overwrite = 0;  // mode 4: SOL = ... instead of SOL += ...
    // mode 4 of the strip kernel: when set, the result is written here as doubles (old SOL still read
    // from out2): the exit of the single-precision multigrid cycle
    double *out64 = nullptr;
    int ldout64 = 0;
};
typedef SpmmEpilogueT<double> SpmmEpilogue;
constexpr int64_t kProfF32 = 100000;  // added to the column count in the profile shape of single-precision launches
void spmm(lb_ctx *c, const lb_mat *a, const double *x, int ldx, double *y, int ldy, int m, int mode = 0,
          const double *b = nullptr, int ldb = 0, const SpmmEpilogue *epi = nullptr);

// single-precision twin (strip kernel only; values = a lazily made float copy kept with the matrix):
// the multigrid cycle that preconditions the eigensolver runs in fp32 (amg.cu).  All operands must have
// 16-byte aligned rows and m % 4 == 0; spmm_f32_supported(): the matrix fits the strip kernel.
void spmm_f32(lb_ctx *c, const lb_mat *a, const float *x, int ldx, float *y, int ldy, int m, int mode = 0,
              const float *b = nullptr, int ldb = 0, const SpmmEpilogueT<float> *epi = nullptr);
bool spmm_f32_supported(lb_ctx *c, const lb_mat *a);
const float *mat_values_f32(lb_ctx *c, const lb_mat *a);
extern thread_local int g_spmm_force_rowwise, g_spmm_variant;
// kernel-level parity of the SpMM forms (lb_spmm_selftest): see blockvec.cu
void spmm_selftest(lb_ctx *c, const lb_mat *a, int m, double *errs);

// out[j] = sum_i X[i,j] * Y[i,j], j < cols  (deterministic two-stage reduction), device output
void col_dots(lb_ctx *c, int64_t n, int cols, const double *x, int ldx, const double *y, int ldy, double *out);

// out[j] = ||AX[:,j] - lam[j] BX[:,j]||^2, out[cols + j] = ||BX[:,j]||^2 (one pass, deterministic)
void residual_norms(lb_ctx *c, int64_t n, int cols, const double *lam, const double *ax, int ldax, const double *bx,
                    int ldbx, double *out);

// Y[:,j] = a[j]*X[:,j] + b[j]*Y[:,j]; a / b device arrays or NULL (then a_const / b_const)
void axpby_cols(lb_ctx *c, int64_t n, int cols, const double *a, double a_const, const double *x, int ldx,
                const double *b, double b_const, double *y, int ldy);

// Y[:, j] = X[:, idx[j]] - lam[idx[j]] * Z[:, idx[j]]   (compacting gather of active columns)
void residual_cols(lb_ctx *c, int64_t n, int ncols, const int *idx, const double *lam, const double *ax, int ldax,
                   const double *mx, int ldmx, double *out, int ldout);

// the same in single precision, columns padded with zeros to a multiple of 4 (ldout % 4 == 0)
void residual_cols_f32(lb_ctx *c, int64_t n, int ncols, const int *idx, const double *lam, const double *ax, int ldax,
                       const double *mx, int ldmx, float *out, int ldout);

void copy_cols(lb_ctx *c, int64_t n, int cols, const double *x, int ldx, double *y, int ldy);
void scale_rows(lb_ctx *c, int64_t n, int cols, const double *d, const double *x, int ldx, double *y, int ldy);
// subtract from every column its mean (constant null-space projection, lapy/diffgeo.py:156 solve)
void remove_col_means(lb_ctx *c, int64_t n, int cols, double *x, int ldx);
// row0: global index of the first local row (row-partitioned mode: same values as one process)
void fill_random(lb_ctx *c, int64_t n, int cols, double *x, int ldx, uint64_t seed, int64_t row0 = 0);
void extract_diagonal(lb_ctx *c, const lb_mat *a, double *d);  // d[i] = A[i,i] (0 if not stored)

// ---- dense tall-skinny products (row-major blocks) -----------------------------------------
// C(p,q) row-major = X(n,p)^T Y(n,q)
// symmetric: the caller guarantees X^T Y is symmetric and p == q (S^T (A S), W^T (B W)): only the
// upper 64x64 tiles are computed and mirrored
void gram(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy, double *cmat,
          bool symmetric = false);
// Y(n,q) = alpha * X(n,p) C(p,q) + beta * Y;  X must not alias Y
void update(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *cmat, int ldc, double alpha,
            double beta, double *y, int ldy);

// hand-written DMMA kernels (dmma.cu); gram()/update() always dispatch to them
void gram_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy, double *cmat,
               bool symmetric);
void update_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *cmat, int ldc,
                 double alpha, double beta, double *y, int ldy);

// X <- X C in place (p <= 64, aligned operands); false: not possible, nothing done
bool update_dmma_inplace(lb_ctx *c, int64_t n, int p, double *x, int ldx, const double *cmat, int ldc);

// benchmark aid (lb_dense_benchmark): ms per launch of op 0 (Gram) / 1 (update), TFLOP/s for op 2 (DMMA peak probe)
double dense_benchmark(lb_ctx *c, int64_t n, int p, int q, int op, int variant, int reps);

// ---- small dense (device, cuSOLVER) ------------------------------------------------------------
// in-place lower Cholesky of the row-major (q,q) SPD matrix g; returns LAPACK info (0 = ok)
int chol_lower(lb_ctx *c, int q, double *g);
// eigen-decomposition of the symmetric row-major (s,s) matrix g: on return row j of g holds the
// j-th eigenvector (ascending), evals (s) device
int sym_eig(lb_ctx *c, int s, double *g, double *evals);
// solve G X = B for SPD row-major G (q,q), B (q, nrhs) row-major ... used for the coarsest level:
// stores the Cholesky factor in g
void dense_chol_solve_prepare(lb_ctx *c, int q, double *g);
// X(q,m) row-major <- G^-1 X with the factor l from dense_chol_solve_prepare
void dense_chol_solve(lb_ctx *c, int q, const double *l, int m, double *x, int ldx);

}  // namespace lb
