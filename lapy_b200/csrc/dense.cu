// dense.cu - dense products on tall-skinny blocks and the small dense factorizations of the
// Rayleigh-Ritz step.
//
// The tall-skinny products go to the hand-written fp64 DMMA kernels (dmma.cu), always; cuSOLVER does
// the <= 3m x 3m Cholesky and symmetric eigenproblem and cuBLAS the triangular solves of the coarsest
// AMG level (SURVEY.md §7 K8 allows a library here: it replaces LAPACK inside ARPACK, is O(m^3) and
// independent of the mesh size).
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cstdlib>
#include <cstring>

#include "blockvec.cuh"

namespace lb {

#define LB_CUBLAS(expr)                                                                        \
    do {                                                                                       \
        cublasStatus_t _s = (expr);                                                            \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                                     \
            set_error("cuBLAS error %d at %s:%d", (int)_s, __FILE__, __LINE__);                \
            throw Error{LB_ERR_CUDA};                                                          \
        }                                                                                      \
    } while (0)
#define LB_CUSOLVER(expr)                                                                      \
    do {                                                                                       \
        cusolverStatus_t _s = (expr);                                                          \
        if (_s != CUSOLVER_STATUS_SUCCESS) {                                                   \
            set_error("cuSOLVER error %d at %s:%d", (int)_s, __FILE__, __LINE__);              \
            throw Error{LB_ERR_CUDA};                                                          \
        }                                                                                      \
    } while (0)

static cublasHandle_t blas(lb_ctx *c) {
    if (!c->cublas) {
        cublasHandle_t h;
        LB_CUBLAS(cublasCreate(&h));
        LB_CUBLAS(cublasSetStream(h, c->stream));
        LB_CUBLAS(cublasSetPointerMode(h, CUBLAS_POINTER_MODE_HOST));
        c->cublas = h;
    }
    return (cublasHandle_t)c->cublas;
}

static cusolverDnHandle_t solver(lb_ctx *c) {
    if (!c->cusolver) {
        cusolverDnHandle_t h;
        LB_CUSOLVER(cusolverDnCreate(&h));
        LB_CUSOLVER(cusolverDnSetStream(h, c->stream));
        c->cusolver = h;
    }
    return (cusolverDnHandle_t)c->cusolver;
}

void destroy_dense_handles(lb_ctx *c) {
    if (c->cublas) cublasDestroy((cublasHandle_t)c->cublas);
    if (c->cusolver) cusolverDnDestroy((cusolverDnHandle_t)c->cusolver);
    c->cublas = c->cusolver = nullptr;
}

void gram(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy, double *cmat,
          bool symmetric) {
    if (p == 0 || q == 0) return;
    ProfScope prof(c, PROF_GRAM, 2.0 * n * p * q, p, q);
    gram_dmma(c, n, p, x, ldx, q, y, ldy, cmat, symmetric && p == q);
}

void update(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *cmat, int ldc, double alpha,
            double beta, double *y, int ldy) {
    if (q == 0 || n == 0) return;
    ProfScope prof(c, PROF_UPDATE, 2.0 * n * p * q, p, q);
    update_dmma(c, n, p, x, ldx, q, cmat, ldc, alpha, beta, y, ldy);
}

int chol_lower(lb_ctx *c, int q, double *g) {
    int lwork = 0;
    LB_CUSOLVER(cusolverDnDpotrf_bufferSize(solver(c), CUBLAS_FILL_MODE_UPPER, q, g, q, &lwork));
    DBuf<double> work(c, lwork);
    DBuf<int> info(c, 1);
    // column-major upper U with G = U^T U is the row-major lower L with G = L L^T
    LB_CUSOLVER(cusolverDnDpotrf(solver(c), CUBLAS_FILL_MODE_UPPER, q, g, q, work.p, lwork, info.p));
    c->launches++;
    int h = 0;
    read_back(c, &h, info.p, 1);
    return h;
}

int sym_eig(lb_ctx *c, int s, double *g, double *evals) {
    ProfScope prof(c, PROF_TRSM, 9.0 * s * s * s, s, 0);  // reported as class "small_dense" (syevd / coarse solves)
    int lwork = 0;
    LB_CUSOLVER(cusolverDnDsyevd_bufferSize(solver(c), CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, s, g, s, evals,
                                            &lwork));
    DBuf<double> work(c, lwork);
    DBuf<int> info(c, 1);
    LB_CUSOLVER(cusolverDnDsyevd(solver(c), CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, s, g, s, evals, work.p,
                                 lwork, info.p));
    c->launches++;
    int h = 0;
    read_back(c, &h, info.p, 1);
    return h;  // eigenvector j is column j column-major == row j row-major
}

void dense_chol_solve_prepare(lb_ctx *c, int q, double *g) {
    int info = chol_lower(c, q, g);
    if (info != 0) {
        set_error("coarsest-level Cholesky failed (info=%d): matrix not positive definite", info);
        throw Error{LB_ERR_NOCONV};
    }
}

// X_rm(q,m) <- G^-1 X given the factor from dense_chol_solve_prepare (row-major lower L)
void dense_chol_solve(lb_ctx *c, int q, const double *l, int m, double *x, int ldx) {
    ProfScope prof(c, PROF_TRSM, 2.0 * q * q * m, q, m);
    // row-major X(q,m) is column-major X^T (m,q): solve X^T <- X^T G^-1 = X^T (U^T U)^-1 from the right
    const double one = 1.0;
    LB_CUBLAS(cublasDtrsm(blas(c), CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, q,
                          &one, l, q, x, ldx));
    LB_CUBLAS(cublasDtrsm(blas(c), CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m, q,
                          &one, l, q, x, ldx));
    c->launches += 2;
}

}  // namespace lb
