// dmma.cu - hand-written fp64 tensor-core (DMMA, mma.sync.m8n8k4.f64) kernels for the two dense
// tall-skinny contractions of the Rayleigh-Ritz step:
//     Gram    C(p,q) = X(n,p)^T Y(n,q)          reduction over the n ~ 10^6 rows, split-K + fixed-order reduce
//     update  Y(n,q) = alpha X(n,p) C(p,q) + beta Y
// They replace the LAPACK/BLAS calls inside ARPACK (eigsh at lapy/solver.py:713).  fp64 tensor work
// on sm_100a is mma.sync DMMA (tcgen05 has no f64 kind); operands are row-major block vectors, so
// fragments are gathered straight from global memory in 32/64-byte row pieces (sector-exact) and
// only the small coefficient matrix is staged in shared memory.
//
// Fragment layout of mma.m8n8k4 (lane = 4*g + t): A[g][t], B[t][g], D[g][2t], D[g][2t+1].
// The k index inside an 8-chunk is permuted (lane t owns k = 2t and 2t+1, used by two successive
// MMAs) so that the A operand of `update` is one 16-byte load per lane; a sum over k is
// order-independent up to rounding and the order is fixed -> bit-reproducible.
#include <cstdlib>

#include "blockvec.cuh"

namespace lb {

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// update: CTA = 8 warps = 128 rows x 64 columns; warp = 32 x 32; persistent over row tiles.
// ---------------------------------------------------------------------------------------------
constexpr int kUpdStride = 66;  // smem row stride of the coefficient slice: conflict-free B frags

__global__ void __launch_bounds__(256, 2)
    update_dmma_kernel(int64_t n, int p, const double *__restrict__ x, int ldx, int q, const double *__restrict__ cmat,
                       int ldc, double alpha, double beta, double *y, int ldy) {
    extern __shared__ double cs[];  // (p8, kUpdStride): C[:, col tile], zero padded
    const int p8 = (p + 7) & ~7;
    const int col_tile = blockIdx.y * 64;
    for (int i = threadIdx.x; i < p8 * 64; i += blockDim.x) {
        const int k = i >> 6, cc = i & 63;
        const int col = col_tile + cc;
        cs[k * kUpdStride + cc] = (k < p && col < q) ? cmat[(int64_t)k * ldc + col] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const bool vec_ok = (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t ntiles = (n + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * 128 + wr;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        const double *xr[4];
        bool rok[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = r0 + 8 * i + g;
            rok[i] = r < n;
            xr[i] = x + (rok[i] ? r : 0) * ldx;
        }
        for (int k8 = 0; k8 < p8; k8 += 8) {
            const int ka = k8 + 2 * t;
            double a0[4], a1[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (rok[i] && ka + 1 < p && vec_ok) {
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(xr[i] + ka));
                    a0[i] = v.x;
                    a1[i] = v.y;
                } else {
                    a0[i] = (rok[i] && ka < p) ? __ldg(xr[i] + ka) : 0.0;
                    a1[i] = (rok[i] && ka + 1 < p) ? __ldg(xr[i] + ka + 1) : 0.0;
                }
            }
            double b0[4], b1[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                b0[j] = cs[ka * kUpdStride + wc + 8 * j + g];
                b1[j] = cs[(ka + 1) * kUpdStride + wc + 8 * j + g];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
                    dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = r0 + 8 * i + g;
            if (r >= n) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = col_tile + wc + 8 * j + 2 * t;
                double *yp = y + r * ldy + col;
                if (col < q) yp[0] = beta == 0.0 ? alpha * acc[i][j][0] : alpha * acc[i][j][0] + beta * yp[0];
                if (col + 1 < q) yp[1] = beta == 0.0 ? alpha * acc[i][j][1] : alpha * acc[i][j][1] + beta * yp[1];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// update, pipelined form: the same 128 x 64 CTA tile, but X (128 rows x 8 k) and C (8 k x 64 cols)
// chunks stream through a 4-stage cp.async ring in shared memory, so the tensor pipe is not
// stalled on global-load latency (v1: 61 % tensor-active).  Needs 16-byte aligned rows of X and C;
// the dispatcher falls back to the direct-load kernel otherwise.
//   Xs[stage][row][8]      stride 8 doubles: quarter-warps read 8 x 16 B conflict-free
//   Cs[stage][k][kUpdStride]
// ---------------------------------------------------------------------------------------------
constexpr int kUpdStages = 4;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__global__ void __launch_bounds__(256, 2)
    update_dmma_pipe_kernel(int64_t n, int p, const double *__restrict__ x, int ldx, int q,
                            const double *__restrict__ cmat, int ldc, double alpha, double beta, double *y, int ldy) {
    extern __shared__ __align__(16) double smem_pipe[];
    double *xs = smem_pipe;                                 // [stages][128][8]
    double *cs = smem_pipe + kUpdStages * 128 * 8;          // [stages][8][kUpdStride]
    const int p8 = (p + 7) & ~7, nchunk = p8 / 8;
    const int col_tile = blockIdx.y * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int64_t ntiles = (n + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row_base = tile * 128;
        auto issue = [&](int chunk, int stage) {
            const int k8 = chunk * 8;
            // X chunk: 128 rows x 64 B = 512 pieces of 16 B, two per thread
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int piece = threadIdx.x + it * 256;
                const int r = piece >> 2, part = piece & 3;
                const int64_t row = row_base + r;
                const int k = k8 + 2 * part;
                const bool ok = row < n && k < p;
                const double *src = x + (ok ? row : 0) * ldx + (ok ? k : 0);
                double *dst = xs + ((size_t)stage * 128 + r) * 8 + 2 * part;
                if (ok && k + 1 >= p) {  // last odd column: plain stores of the single valid value
                    dst[0] = __ldg(src);
                    dst[1] = 0.0;
                } else {
                    cp_async16(dst, src, ok);
                }
            }
            // C chunk: 8 rows x 64 cols = 256 pieces of 16 B, one per thread
            {
                const int r = threadIdx.x >> 5, part = threadIdx.x & 31;
                const int k = k8 + r, col = col_tile + 2 * part;
                const bool ok = k < p && col + 1 < q;
                const double *src = cmat + (int64_t)(k < p ? k : 0) * ldc + (col + 1 < q ? col : 0);
                double *dst = cs + ((size_t)stage * 8 + r) * kUpdStride + 2 * part;
                if (!ok && k < p && col < q) {  // ragged last column
                    dst[0] = __ldg(cmat + (int64_t)k * ldc + col);
                    dst[1] = 0.0;
                } else {
                    cp_async16(dst, src, ok);
                }
            }
        };
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        __syncthreads();  // the previous tile's readers are done with the ring
#pragma unroll
        for (int s0 = 0; s0 < kUpdStages - 1; s0++) {
            if (s0 < nchunk) issue(s0, s0);
            cp_async_commit();
        }
        for (int ch = 0; ch < nchunk; ch++) {
            cp_async_wait<kUpdStages - 2>();
            __syncthreads();  // chunk ch has landed for every thread; stage (ch-1) is free again
            const int nx = ch + kUpdStages - 1;
            if (nx < nchunk) issue(nx, nx % kUpdStages);
            cp_async_commit();
            const int stage = ch % kUpdStages;
            const double *xst = xs + (size_t)stage * 128 * 8;
            const double *cst = cs + (size_t)stage * 8 * kUpdStride;
            double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double2 v = *reinterpret_cast<const double2 *>(xst + (wr + 8 * i + g) * 8 + 2 * t);
                a0[i] = v.x;
                a1[i] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                b0[j] = cst[(2 * t) * kUpdStride + wc + 8 * j + g];
                b1[j] = cst[(2 * t + 1) * kUpdStride + wc + 8 * j + g];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
                    dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
                }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = row_base + wr + 8 * i + g;
            if (r >= n) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = col_tile + wc + 8 * j + 2 * t;
                double *yp = y + r * ldy + col;
                if (col < q) yp[0] = beta == 0.0 ? alpha * acc[i][j][0] : alpha * acc[i][j][0] + beta * yp[0];
                if (col + 1 < q) yp[1] = beta == 0.0 ? alpha * acc[i][j][1] : alpha * acc[i][j][1] + beta * yp[1];
            }
        }
    }
}

void update_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *cmat, int ldc,
                 double alpha, double beta, double *y, int ldy) {
    static int use_pipe = -1;
    if (use_pipe < 0) {
        const char *e = getenv("LAPY_B200_UPDATE");
        use_pipe = (e && !strcmp(e, "direct")) ? 0 : 1;
    }
    const bool aligned = (ldx % 2 == 0) && (ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(cmat) & 15) == 0);
    if (use_pipe && aligned) {
        const size_t smem = (size_t)kUpdStages * (128 * 8 + 8 * kUpdStride) * sizeof(double);
        // per device (a process may drive several contexts): set every time, it is a cheap host call
        LB_CUDA(cudaFuncSetAttribute(update_dmma_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int ytiles = cdiv(q, 64);
        const int64_t ntiles = (n + 127) / 128;
        const int gx = (int)std::min<int64_t>(ntiles, std::max(1, (kSMs * 2) / ytiles));
        dim3 grid(gx, ytiles);
        LB_LAUNCH(c, update_dmma_pipe_kernel, grid, 256, smem, n, p, x, ldx, q, cmat, ldc, alpha, beta, y, ldy);
        return;
    }
    const int p8 = (p + 7) & ~7;
    const size_t smem = (size_t)p8 * kUpdStride * sizeof(double);
    LB_CUDA(cudaFuncSetAttribute(update_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LB_REQUIRE(smem <= 200 * 1024, "block update: inner dimension %d too large", p);
    const int ytiles = cdiv(q, 64);
    const int64_t ntiles = (n + 127) / 128;
    const int gx = (int)std::min<int64_t>(ntiles, std::max(1, (kSMs * 2) / ytiles));
    dim3 grid(gx, ytiles);
    LB_LAUNCH(c, update_dmma_kernel, grid, 256, smem, n, p, x, ldx, q, cmat, ldc, alpha, beta, y, ldy);
}

// ---------------------------------------------------------------------------------------------
// Gram: CTA = 4 warps = one 64 x 64 output tile over a slab of rows; warp = 32 x 32.
// partial[(split * ntile + tile) * 4096 + i*64 + j]; gram_reduce sums the slabs in fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4)
    gram_dmma_kernel(int64_t n, int p, const double *__restrict__ x, int ldx, int q, const double *__restrict__ y,
                     int ldy, int qtiles, int symmetric, int64_t rows_per_split, double *__restrict__ partial,
                     int ntile_total) {
    const int tile = blockIdx.x;
    const int pt = tile / qtiles, qt = tile % qtiles;
    if (symmetric && qt < pt) return;  // mirrored by gram_reduce
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int pc = pt * 64 + (warp & 1) * 32, qc = qt * 64 + (warp >> 1) * 32;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(n, r_begin + rows_per_split);
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    bool aok[4], bok[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        aok[i] = pc + 8 * i + g < p;
        bok[i] = qc + 8 * i + g < q;
    }
    for (int64_t r8 = r_begin; r8 < r_end; r8 += 8) {
        const int64_t ra = r8 + 2 * t, rb = ra + 1;
        const bool va = ra < r_end, vb = rb < r_end;
        const double *xa = x + ra * ldx + pc + g, *xb = x + rb * ldx + pc + g;
        const double *ya = y + ra * ldy + qc + g, *yb = y + rb * ldy + qc + g;
        double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a0[i] = (va && aok[i]) ? __ldg(xa + 8 * i) : 0.0;
            a1[i] = (vb && aok[i]) ? __ldg(xb + 8 * i) : 0.0;
            b0[i] = (va && bok[i]) ? __ldg(ya + 8 * i) : 0.0;
            b1[i] = (vb && bok[i]) ? __ldg(yb + 8 * i) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
                dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
            }
    }
    double *out = partial + ((int64_t)blockIdx.y * ntile_total + tile) * 4096;
    const int li = (warp & 1) * 32, lj = (warp >> 1) * 32;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2 *>(out + (li + 8 * i + g) * 64 + lj + 8 * j + 2 * t) = v;
        }
}

__global__ void gram_reduce_kernel(int p, int q, int qtiles, int ntile_total, int nsplit, int symmetric,
                                   const double *__restrict__ partial, double *__restrict__ cmat) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * q) return;
    const int i = idx / q, j = idx % q;
    int pi = i, pj = j;
    if (symmetric && (j >> 6) < (i >> 6)) {  // lower tile: take the transposed entry of the upper tile
        pi = j;
        pj = i;
    }
    const int tile = (pi >> 6) * qtiles + (pj >> 6);
    const double *src = partial + (int64_t)tile * 4096 + (pi & 63) * 64 + (pj & 63);
    double s = 0.0;
    for (int k = 0; k < nsplit; k++) s += src[(int64_t)k * ntile_total * 4096];
    cmat[idx] = s;
}

// symmetric != 0: the caller guarantees X^T Y is symmetric (S^T (A S), W^T (B W)) and p == q
void gram_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy, double *cmat,
               bool symmetric) {
    const int ptiles = cdiv(p, 64), qtiles = cdiv(q, 64);
    const int ntile = ptiles * qtiles;
    const int active = symmetric ? ptiles * (ptiles + 1) / 2 : ntile;
    int nsplit = std::max(1, (kSMs * 4) / active);
    int64_t rows_per_split = ((n + nsplit - 1) / nsplit + 7) & ~7ll;
    if (rows_per_split < 512) rows_per_split = 512;
    nsplit = (int)((n + rows_per_split - 1) / rows_per_split);
    DBuf<double> partial(c, (size_t)nsplit * ntile * 4096);
    dim3 grid(ntile, nsplit);
    LB_LAUNCH(c, gram_dmma_kernel, grid, 128, 0, n, p, x, ldx, q, y, ldy, qtiles, (int)symmetric, rows_per_split,
              partial.p, ntile);
    LB_LAUNCH(c, gram_reduce_kernel, cdiv(p * q, 256), 256, 0, p, q, qtiles, ntile, nsplit, (int)symmetric, partial.p,
              cmat);
}

}  // namespace lb
