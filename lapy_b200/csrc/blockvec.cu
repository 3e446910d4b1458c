// blockvec.cu - CSR x block-vector products and block BLAS-1 kernels (all HBM-bound).
//
// Replaces scipy's cs*_matvec(s) (lapy/solver.py:844-846, ARPACK's M/OPinv callbacks) and the
// NumPy vector arithmetic inside ARPACK / SuperLU solves with kernels that keep every block
// vector resident on the device.  Reductions are two-stage with a fixed tree -> bit-reproducible.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "blockvec.cuh"

namespace lb {

// ---- SpMM: a group of G lanes per row, lanes over columns -------------------------------------
// Each nonzero (j, a) is broadcast to the group; the group streams row j of X as one contiguous
// segment (8*m bytes), so X traffic is fully coalesced; matrix entries are read once per row.
// A CTA owns a contiguous strip of kSpmmStrip rows and its warps walk the strip interleaved, so the
// X rows shared by neighbouring matrix rows (mesh neighbours after the locality renumbering) are
// served by the SM's L1 instead of L2: the L2 -> SM traffic drops from ~nnz/row x to ~1-2x |X|.
constexpr int kSpmmStrip = 128;

template <int G, bool VEC>
__global__ void __launch_bounds__(256) spmm_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                                   const int32_t *__restrict__ indices,
                                                   const double *__restrict__ val, const double *__restrict__ x,
                                                   int ldx, double *y, int ldy, int m, int mode,
                                                   const double *b, int ldb, SpmmEpilogue epi) {
    // ncu (round 1): this kernel is L1/TEX-throughput bound (73 %), not DRAM bound (44 %): every
    // nonzero cost one 512-byte X request plus two broadcast requests for (index, value).  The
    // group now fetches a row's (index, value) pairs with ONE coalesced load each (lane q holds
    // entry q) and broadcasts them by warp shuffle, leaving only the X gathers on the L1 pipe.
    // (Staging the strip's CSR segment in shared memory instead was measured 14 % slower.)
    // Round 2, ncu source view: 302 warp instructions per row, stall samples on the first shuffle (index
    // load) and the first fma (X rows) = three dependent latencies per row.  Software pipelining the
    // row loop (row pointers two rows ahead, entries one row ahead) was measured SLOWER (1.25 vs 1.11 ms
    // at 64 columns, every width): the extra live registers cost more occupancy than the overlap wins.
    constexpr int GROUPS = 256 / G;  // rows in flight per CTA
    const int grp = threadIdx.x / G, lane = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int64_t strip0 = (int64_t)blockIdx.x * kSpmmStrip;
    const int nrows = (int)(min(n, strip0 + kSpmmStrip) - strip0);
    for (int lr = grp; lr < nrows; lr += GROUPS) {
        const int64_t row = strip0 + lr;
        const int beg = __ldg(indptr + row), end = __ldg(indptr + row + 1);
        // VEC: lane owns columns c0 + 2*lane, +1 (one 16-byte load); else c0 + lane, c0 + lane + G
        for (int c0 = 0; c0 < m; c0 += 2 * G) {
            const int ca = VEC ? c0 + 2 * lane : c0 + lane;
            const int cb = VEC ? ca + 1 : ca + G;
            const bool ha = ca < m, hb = cb < m;
            double s0 = 0.0, s1 = 0.0;
            for (int p0 = beg; p0 < end; p0 += G) {
                const int cnt = min(G, end - p0);
                int jl = 0;
                double al = 0.0;
                if (lane < cnt) {
                    jl = __ldg(indices + p0 + lane);
                    al = __ldg(val + p0 + lane);
                }
                int q = 0;
                for (; q + 3 < cnt; q += 4) {
                    int j[4];
                    double a[4], u0[4], u1[4];
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        j[w] = __shfl_sync(gmask, jl, q + w, G);
                        a[w] = __shfl_sync(gmask, al, q + w, G);
                    }
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const double *xr = x + (int64_t)j[w] * ldx;
                        if (VEC && hb) {
                            const double2 v = __ldg(reinterpret_cast<const double2 *>(xr + ca));
                            u0[w] = v.x;
                            u1[w] = v.y;
                        } else {
                            u0[w] = ha ? __ldg(xr + ca) : 0.0;
                            u1[w] = hb ? __ldg(xr + cb) : 0.0;
                        }
                    }
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        s0 = fma(a[w], u0[w], s0);
                        s1 = fma(a[w], u1[w], s1);
                    }
                }
                for (; q < cnt; q++) {
                    const int j0 = __shfl_sync(gmask, jl, q, G);
                    const double a0 = __shfl_sync(gmask, al, q, G);
                    const double *xr = x + (int64_t)j0 * ldx;
                    if (VEC && hb) {
                        const double2 v = __ldg(reinterpret_cast<const double2 *>(xr + ca));
                        s0 = fma(a0, v.x, s0);
                        s1 = fma(a0, v.y, s1);
                    } else {
                        if (ha) s0 = fma(a0, __ldg(xr + ca), s0);
                        if (hb) s1 = fma(a0, __ldg(xr + cb), s1);
                    }
                }
            }
            if (mode == 1) {
                if (ha) s0 = b[row * ldb + ca] - s0;
                if (hb) s1 = b[row * ldb + cb] - s1;
            } else if (mode == 2) {
                if (ha) s0 = b[row * ldb + ca] + s0;
                if (hb) s1 = b[row * ldb + cb] + s1;
            } else if (mode == 3) {
                // fused first Chebyshev step: r = b - K x (written to y), d = c2 * dinv o r (out2)
                const double di = epi.c2 * __ldg(epi.dinv + row);
                if (ha) {
                    s0 = b[row * ldb + ca] - s0;
                    epi.out2[row * epi.ldout2 + ca] = di * s0;
                }
                if (hb) {
                    s1 = b[row * ldb + cb] - s1;
                    epi.out2[row * epi.ldout2 + cb] = di * s1;
                }
            } else if (mode == 4) {
                // fused last Chebyshev step (x gathers d): rr = b - K d; dn = c1 d + c2 dinv o rr;
                // sol (+)= d + dn; neither the residual nor dn is written
                const double di = epi.c2 * __ldg(epi.dinv + row);
                if (ha) {
                    const double dold = x[row * ldx + ca];
                    const double dn = fma(epi.c1, dold, di * (b[row * ldb + ca] - s0));
                    double *sp = epi.out2 + row * epi.ldout2 + ca;
                    *sp = (epi.overwrite ? 0.0 : *sp) + dold + dn;
                }
                if (hb) {
                    const double dold = x[row * ldx + cb];
                    const double dn = fma(epi.c1, dold, di * (b[row * ldb + cb] - s1));
                    double *sp = epi.out2 + row * epi.ldout2 + cb;
                    *sp = (epi.overwrite ? 0.0 : *sp) + dold + dn;
                }
                continue;
            }
            if (ha) y[row * ldy + ca] = s0;
            if (hb) y[row * ldy + cb] = s1;
        }
    }
}

// ---- SpMV, CSR-stream form: the CTA stages the products a_ij * x_j of a strip of rows in shared
// memory (one independent gather per thread and entry: maximal memory-level parallelism, CSR
// arrays read fully coalesced), then one thread per row sums its segment left to right.
// The CSR arrays are touched exactly once, so they are loaded with the streaming (evict-first)
// hint and leave L1/L2 to the gathered x: measured on the level-9 operator 55 % -> 68 % of the
// HBM roofline in the solver numbering, 46 % -> 58 % in the caller's (tools/sweep_spmv.py).
constexpr int kSpmvCap = 2944;  // 2 x 23 KB products + row pointers < 48 KB static
constexpr int kSpmvBatch = 2;   // (index, value) pairs loaded per thread before the dependent gathers

// ROWS per CTA: 256 (short rows, ~7 entries) or 128 (tets, ~15)
template <int MC, int ROWS>
__global__ void __launch_bounds__(256) spmv_stream_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                                          const int32_t *__restrict__ indices,
                                                          const double *__restrict__ val,
                                                          const double *__restrict__ x, int ldx, double *y, int ldy,
                                                          int m, int mode, const double *b, int ldb) {
    __shared__ double s_prod[MC][kSpmvCap];
    __shared__ int32_t s_ptr[ROWS + 1];
    const int64_t strip0 = (int64_t)blockIdx.x * ROWS;
    const int nrows = (int)(min(n, strip0 + ROWS) - strip0);
    for (int i = threadIdx.x; i <= nrows; i += 256) s_ptr[i] = __ldg(indptr + strip0 + i);
    __syncthreads();
    const int base = s_ptr[0], total = s_ptr[nrows] - base;
    if (total <= kSpmvCap) {
        for (int i0 = 0; i0 < total; i0 += 256 * kSpmvBatch) {
            int j[kSpmvBatch];
            double a[kSpmvBatch];
#pragma unroll
            for (int u = 0; u < kSpmvBatch; u++) {
                const int i = i0 + u * 256 + threadIdx.x;
                j[u] = i < total ? __ldcs(indices + base + i) : 0;
                a[u] = i < total ? __ldcs(val + base + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kSpmvBatch; u++) {
                const int i = i0 + u * 256 + threadIdx.x;
                if (i < total) {
                    s_prod[0][i] = a[u] * __ldg(x + (int64_t)j[u] * ldx);
                    if (MC > 1) s_prod[MC - 1][i] = a[u] * __ldg(x + (int64_t)j[u] * ldx + 1);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < nrows) {
            const int64_t row = strip0 + threadIdx.x;
            const int beg = s_ptr[threadIdx.x] - base, end = s_ptr[threadIdx.x + 1] - base;
            for (int col = 0; col < m; col++) {
                double s = 0.0;
                for (int p = beg; p < end; p++) s += s_prod[col][p];
                if (mode == 1) s = b[row * ldb + col] - s;
                else if (mode == 2) s = b[row * ldb + col] + s;
                y[row * ldy + col] = s;
            }
        }
    } else {  // very long rows: one thread per row straight from global memory
        if (threadIdx.x < nrows) {
            const int64_t row = strip0 + threadIdx.x;
            for (int col = 0; col < m; col++) {
                double s = 0.0;
                for (int p = s_ptr[threadIdx.x]; p < s_ptr[threadIdx.x + 1]; p++)
                    s += __ldg(val + p) * __ldg(x + (int64_t)__ldg(indices + p) * ldx + col);
                if (mode == 1) s = b[row * ldb + col] - s;
                else if (mode == 2) s = b[row * ldb + col] + s;
                y[row * ldy + col] = s;
            }
        }
    }
}

// diagonal matrix (lumped mass, identity): entries only on the diagonal, possibly missing rows
__global__ void diag_spmm_kernel(int64_t n, const int32_t *__restrict__ indptr, const double *__restrict__ val,
                                 const double *__restrict__ x, int ldx, double *y, int ldy, int m,
                                 int mode, const double *b, int ldb) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t row = t / m;
    const int col = (int)(t - row * m);
    const int beg = indptr[row], end = indptr[row + 1];
    double s = end > beg ? val[beg] * x[row * ldx + col] : 0.0;
    if (mode == 1) s = b[row * ldb + col] - s;
    else if (mode == 2) s = b[row * ldb + col] + s;
    y[row * ldy + col] = s;
}

void spmm(lb_ctx *c, const lb_mat *a, const double *x, int ldx, double *y, int ldy, int m, int mode, const double *b,
          int ldb, const SpmmEpilogue *epi_in) {
    SpmmEpilogue epi{};
    if (epi_in) epi = *epi_in;
    LB_REQUIRE(mode <= 2 || (epi_in && !a->diagonal && m > 2), "fused SpMM epilogues need a general matrix and m > 2");
    const int64_t n = a->n;
    if (n == 0 || m == 0) return;
    const int64_t xrows = a->ncols < 0 ? a->n : a->ncols;
    // algorithmic bytes: matrix once, X once, Y once (+ B once for the residual modes)
    ProfScope prof(c, PROF_SPMM, 12.0 * a->nnz + 4.0 * (n + 1) + 8.0 * m * (xrows + n * (mode ? 2 : 1)), m, a->nnz);
    if (a->diagonal) {
        LB_LAUNCH(c, diag_spmm_kernel, cdiv(n * m, 256), 256, 0, n, a->indptr.p, a->data.p, x, ldx, y, ldy, m, mode, b,
                  ldb);
        return;
    }
    const int32_t *ip = a->indptr.p, *ix = a->indices.p;
    const double *v = a->data.p;
    if (m <= 2) {
        // 256 rows per CTA when the strip's entries fit the staging buffer on average (<= 10 per row)
        const bool wide = a->nnz <= 10 * n;
#define LB_SPMV(MC, ROWS) LB_LAUNCH(c, (spmv_stream_kernel<MC, ROWS>), cdiv(n, ROWS), 256, 0, n, ip, ix, v, x, ldx, y, ldy, m, mode, b, ldb)
        if (m == 1 && wide) LB_SPMV(1, 256);
        else if (m == 1) LB_SPMV(1, 128);
        else if (wide) LB_SPMV(2, 256);
        else LB_SPMV(2, 128);
#undef LB_SPMV
        return;
    }
    // 16-byte vector loads of X need even leading dimension and a 16-byte aligned base
    const bool vec = (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int grid = cdiv(n, kSpmmStrip);
    // (Round 2 tried a variant that copies the strip's own 128 rows of X into shared memory with
    // cp.async.bulk + mbarrier and serves the in-strip gathers with LDS.128: 1.45 ms instead of 1.11 ms
    // at 64 columns, slower at every width - the copy latency is exposed at the head of every CTA and
    // 64 KB of shared memory leave 3 CTAs per SM.  Removed; profiles/spmm_shapes_r2.json has both.)
#define LB_SPMM(G)                                                                                       \
    do {                                                                                                 \
        if (vec) LB_LAUNCH(c, (spmm_kernel<G, true>), grid, 256, 0, n, ip, ix, v, x, ldx, y, ldy, m, mode, b, ldb, epi); \
        else LB_LAUNCH(c, (spmm_kernel<G, false>), grid, 256, 0, n, ip, ix, v, x, ldx, y, ldy, m, mode, b, ldb, epi);    \
    } while (0)
    if (m <= 8) LB_SPMM(4);
    else if (m <= 16) LB_SPMM(8);
    else if (m <= 32) LB_SPMM(16);
    else LB_SPMM(32);
#undef LB_SPMM
}

// ---- column dots ------------------------------------------------------------------------------
constexpr int kDotBlocks = kSMs * 4;
constexpr int kDotCW = 32;  // columns per pass
constexpr int kDotRY = 8;   // row lanes per block (block = 32 x 8 threads)

__global__ void __launch_bounds__(kDotCW *kDotRY) col_dots_partial(int64_t n, int cols, const double *__restrict__ x,
                                                                   int ldx, const double *__restrict__ y, int ldy,
                                                                   double *__restrict__ partial) {
    __shared__ double red[kDotRY][kDotCW + 1];
    const int tx = threadIdx.x % kDotCW, ty = threadIdx.x / kDotCW;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
    for (int c0 = 0; c0 < cols; c0 += kDotCW) {
        const int col = c0 + tx;
        double s = 0.0;
        if (col < cols)
            for (int64_t r = r0 + ty; r < r1; r += kDotRY) s = fma(x[r * ldx + col], y ? y[r * ldy + col] : 1.0, s);
        red[ty][tx] = s;
        __syncthreads();
        if (ty == 0 && col < cols) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < kDotRY; k++) t += red[k][tx];
            partial[(int64_t)blockIdx.x * cols + col] = t;
        }
        __syncthreads();
    }
}

__global__ void col_dots_final(int nblocks, int cols, const double *__restrict__ partial, double *__restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(int64_t)b * cols + col];
    out[col] = s;
}

void col_dots(lb_ctx *c, int64_t n, int cols, const double *x, int ldx, const double *y, int ldy, double *out) {
    if (cols == 0) return;
    ProfScope prof(c, PROF_DOTS, 8.0 * n * cols * (y && y != x ? 2 : 1));
    const int nb = (int)std::min<int64_t>(kDotBlocks, std::max<int64_t>(1, n / 64));
    DBuf<double> partial(c, (size_t)nb * cols);
    LB_LAUNCH(c, col_dots_partial, nb, kDotCW * kDotRY, 0, n, cols, x, ldx, y, ldy, partial.p);
    LB_LAUNCH(c, col_dots_final, cdiv(cols, 64), 64, 0, nb, cols, partial.p, out);
}

// ---- elementwise block kernels -------------------------------------------------------------------
__global__ void axpby_cols_kernel(int64_t n, int cols, const double *__restrict__ a, double a_const,
                                  const double *__restrict__ x, int ldx, const double *__restrict__ b, double b_const,
                                  double *__restrict__ y, int ldy) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t row = t / cols;
    const int col = (int)(t - row * cols);
    const double av = a ? a[col] : a_const, bv = b ? b[col] : b_const;
    const double xv = x[row * ldx + col];
    double *yp = y + row * ldy + col;
    *yp = bv == 0.0 ? av * xv : fma(av, xv, bv * *yp);
}

void axpby_cols(lb_ctx *c, int64_t n, int cols, const double *a, double a_const, const double *x, int ldx,
                const double *b, double b_const, double *y, int ldy) {
    if (n * cols == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 24.0 * n * cols);
    LB_LAUNCH(c, axpby_cols_kernel, cdiv(n * cols, 256), 256, 0, n, cols, a, a_const, x, ldx, b, b_const, y, ldy);
}

__global__ void residual_cols_kernel(int64_t n, int ncols, const int *__restrict__ idx, const double *__restrict__ lam,
                                     const double *__restrict__ ax, int ldax, const double *__restrict__ mx, int ldmx,
                                     double *__restrict__ out, int ldout) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ncols) return;
    const int64_t row = t / ncols;
    const int a = (int)(t - row * ncols);
    const int col = idx[a];
    out[row * ldout + a] = fma(-lam[col], mx[row * ldmx + col], ax[row * ldax + col]);
}

void residual_cols(lb_ctx *c, int64_t n, int ncols, const int *idx, const double *lam, const double *ax, int ldax,
                   const double *mx, int ldmx, double *out, int ldout) {
    if (n * ncols == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 24.0 * n * ncols);
    LB_LAUNCH(c, residual_cols_kernel, cdiv(n * ncols, 256), 256, 0, n, ncols, idx, lam, ax, ldax, mx, ldmx, out, ldout);
}

void copy_cols(lb_ctx *c, int64_t n, int cols, const double *x, int ldx, double *y, int ldy) {
    if (n * cols == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 16.0 * n * cols);
    LB_CUDA(cudaMemcpy2DAsync(y, (size_t)ldy * 8, x, (size_t)ldx * 8, (size_t)cols * 8, n, cudaMemcpyDeviceToDevice,
                              c->stream));
}

__global__ void scale_rows_kernel(int64_t n, int cols, const double *__restrict__ d, const double *__restrict__ x,
                                  int ldx, double *__restrict__ y, int ldy) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t row = t / cols;
    const int col = (int)(t - row * cols);
    y[row * ldy + col] = d[row] * x[row * ldx + col];
}

void scale_rows(lb_ctx *c, int64_t n, int cols, const double *d, const double *x, int ldx, double *y, int ldy) {
    if (n * cols == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 16.0 * n * cols);
    LB_LAUNCH(c, scale_rows_kernel, cdiv(n * cols, 256), 256, 0, n, cols, d, x, ldx, y, ldy);
}

__global__ void sub_means_kernel(int64_t n, int cols, const double *__restrict__ sums, double *__restrict__ x,
                                 int ldx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t row = t / cols;
    const int col = (int)(t - row * cols);
    x[row * ldx + col] -= sums[col] / (double)n;
}

void remove_col_means(lb_ctx *c, int64_t n, int cols, double *x, int ldx) {
    if (n * cols == 0) return;
    DBuf<double> sums(c, cols);
    col_dots(c, n, cols, x, ldx, nullptr, 0, sums.p);  // y == NULL: dots with ones = column sums
    LB_LAUNCH(c, sub_means_kernel, cdiv(n * cols, 256), 256, 0, n, cols, sums.p, x, ldx);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_random_kernel(int64_t n, int cols, double *__restrict__ x, int ldx, uint64_t seed, int64_t row0) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t row = t / cols;
    const int col = (int)(t - row * cols);
    const uint64_t h = splitmix64(splitmix64(seed + (uint64_t)col) ^ (uint64_t)(row0 + row));
    x[row * ldx + col] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;  // uniform [-1,1)
}

void fill_random(lb_ctx *c, int64_t n, int cols, double *x, int ldx, uint64_t seed, int64_t row0) {
    if (n * cols == 0) return;
    LB_LAUNCH(c, fill_random_kernel, cdiv(n * cols, 256), 256, 0, n, cols, x, ldx, seed, row0);
}

__global__ void extract_diag_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                    const double *__restrict__ val, double *__restrict__ d) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 0.0;
    for (int p = indptr[r]; p < indptr[r + 1]; p++)
        if (indices[p] == r) s = val[p];
    d[r] = s;
}

void extract_diagonal(lb_ctx *c, const lb_mat *a, double *d) {
    if (a->n == 0) return;
    LB_LAUNCH(c, extract_diag_kernel, cdiv(a->n, 256), 256, 0, a->n, a->indptr.p, a->indices.p, a->data.p, d);
}

}  // namespace lb
