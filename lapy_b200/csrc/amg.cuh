// amg.cuh - smoothed-aggregation AMG hierarchy built and applied on the device.
//
// Replaces the sparse factorisation of the reference (SuperLU splu / CHOLMOD, lapy/solver.py:707,
// :876, lapy/heat.py:226) as the "inverse" used by the eigensolver and the linear solves: one
// V-cycle is the preconditioner of block LOBPCG / block PCG.  Setup is deterministic (hash
// priorities, fixed summation orders) and runs entirely on the GPU.
#pragma once
#include <memory>

#include "blockvec.cuh"

namespace lb {

struct AmgLevel {
    std::unique_ptr<lb_mat> K;   // operator of this level (level 0: a private copy alpha*A + beta*B)
    std::unique_ptr<lb_mat> P;   // (n_l, n_{l+1}) smoothed prolongator, CSR
    std::unique_ptr<lb_mat> R;   // P^T, CSR
    DBuf<double> dinv;           // 1 / diag(K)
    double rho = 2.0;            // upper bound of the spectral radius of D^-1 K (Gershgorin)
    // V-cycle work blocks (n_l, mcap): solution, right-hand side, residual, Chebyshev direction
    DBuf<double> x, b, r, d;
    // single-precision twins (amg_prepare_f32)
    DBuf<float> dinv32, x32, b32, r32, d32;
};

struct Amg {
    lb_ctx *ctx = nullptr;
    std::vector<AmgLevel> levels;
    DBuf<double> coarse_chol;  // dense Cholesky factor of the coarsest operator
    // explicit inverse of the coarsest operator, row-major (coarse_n, coarse_ld) with an even leading
    // dimension: the coarse solve of an aligned block is ONE DMMA block product instead of two
    // latency-bound cuBLAS trsm (0.41 ms -> ~0.05 ms per visit at 561 unknowns x 64 columns)
    DBuf<double> coarse_inv;
    int coarse_ld = 0;
    int coarse_n = 0;
    int cheb_deg = 2;
    int gamma = 2;  // cycle index: 1 = V, 2 = W
    int mcap = 0;  // block width the work arrays are sized for
    DBuf<float> coarse_inv32;  // single-precision copy of coarse_inv, leading dimension coarse_ld32
    int coarse_ld32 = 0;
    int f32_state = 0;  // 0: not prepared, 1: single-precision mirrors ready, -1: not supported for this hierarchy
    double setup_ms = 0;
};

struct AmgOptions {
    double theta = 0.0;      // strength-of-connection threshold (0: every off-diagonal is strong)
    int max_coarse = 2000;   // dense coarse solve at or below this size
    int max_levels = 16;
    int cheb_deg = 2;
    int gamma = 2;  // W-cycle: level-independent convergence with MIS-2 aggregates (measured: 108 -> 57 LOBPCG iterations at level 9)
};

// K is consumed (moved into level 0).  mcap: max number of simultaneous right-hand sides.
std::unique_ptr<Amg> amg_setup(lb_ctx *c, std::unique_ptr<lb_mat> K, int mcap, const AmgOptions &opt);
// z (n_level, m) = cycle(r) starting at `level` (0 = finest); r is not modified
void amg_apply(Amg &amg, const double *r, int ldr, double *z, int ldz, int m, int level = 0);
// The same cycle in single precision (the eigensolver's preconditioner).  amg_prepare_f32 builds the
// float mirrors once and says whether the hierarchy supports it.  r: (n_level, ldr) floats with
// m % 4 == 0 columns (callers pad with zero columns) and ldr % 4 == 0; z: doubles, columns 0..mz-1 valid
// (when ldz >= m the padding columns mz..m-1 of a row of z are overwritten with zeros as well).
constexpr int kF32Cycles = 3;  // multigrid cycles per single-precision application (amg.cu)
bool amg_prepare_f32(Amg &amg);
void amg_apply_f32(Amg &amg, const float *r, int ldr, double *z, int ldz, int m, int mz, int level = 0);
// y(n, ldy) floats = x(n, ldx) doubles, columns m..roundup4(m)-1 zero-filled
void convert_cols_f32(lb_ctx *c, int64_t n, int m, const double *x, int ldx, float *y, int ldy);

// C = alpha*A + beta*B for matrices with identical pattern or diagonal B (new matrix)
std::unique_ptr<lb_mat> mat_axpby(lb_ctx *c, const lb_mat *a, double alpha, const lb_mat *b, double beta);
// general sparse product C = A * B (CSR, deterministic, sorted columns)
std::unique_ptr<lb_mat> spgemm(lb_ctx *c, const lb_mat *a, const lb_mat *b);
std::unique_ptr<lb_mat> transpose(lb_ctx *c, const lb_mat *a);
// P A P^T for the renumbering new -> old = order, old -> new = inv
std::unique_ptr<lb_mat> permute_symmetric(lb_ctx *c, const lb_mat *a, const int32_t *order, const int32_t *inv);
// a matrix stored in a mesh's locality numbering (lb_mat::permuted) as a new matrix in the caller's
// numbering with sorted rows (= the canonical CSC the reference's callers see)
std::unique_ptr<lb_mat> to_caller_order(lb_ctx *c, const lb_mat *a);
// Operands of a solver call brought to ONE numbering: if any of them is stored in a locality
// numbering, the others (uploaded by the user, in the caller's numbering) are permuted into it.
struct MatView {
    const lb_mat *m = nullptr;
    std::unique_ptr<lb_mat> owned;
};
// returns the numbering the views are in (nullptr: the caller's); b may be NULL
std::shared_ptr<lb_order> common_numbering(lb_ctx *c, const lb_mat *a, const lb_mat *b, MatView &va, MatView &vb);
// rows [r0, r1) of a (CSR, global columns)
std::unique_ptr<lb_mat> row_block(lb_ctx *c, const lb_mat *a, int64_t r0, int64_t r1, int64_t ncols);
// y[i, :] = x[map[i], :]
void gather_rows(lb_ctx *c, int64_t n, int cols, const int32_t *map, const double *x, int ldx, double *y, int ldy);
// dinv[i] = d[i] > 0 ? 1/d[i] : 0
void launch_diag_inverse(lb_ctx *c, int64_t n, const double *d, double *dinv);

}  // namespace lb
