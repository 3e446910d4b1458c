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
#include <cstring>

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
            // the two DMMAs of an accumulator are issued a full sweep apart (asm volatile keeps this order):
            // back to back they are a dependent pair and the warp idles for the DMMA latency after every other issue
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
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

// Tile shapes: WCOLS x NJ 8-column blocks per CTA.  <2,4>: 128 rows x 64 columns (warps 4 x 2, the
// wide products); <1,NJ>: 256 rows x 8*NJ columns (warps 8 x 1) for the narrow right-hand sides of
// the late LOBPCG iterations (W -= [X P] G with 8..32 active columns): a 64-wide tile would spend
// most of its DMMAs on padding (measured: q = 24 ran at 10 TFLOP/s, q = 64 at 25).
// (Tried in round 2: several k-chunks per barrier - 2 changed nothing, 4 was 10 % slower.)
template <int WCOLS, int NJ>
__global__ void __launch_bounds__(256, 2)
    update_dmma_pipe_kernel(int64_t n, int p, const double *x, int ldx, int q,  // x may alias y (in place, one column tile)
                            const double *__restrict__ cmat, int ldc, double alpha, double beta, double *y, int ldy) {
    constexpr int ROWS = WCOLS == 2 ? 128 : 256;
    constexpr int COLS = WCOLS * NJ * 8;
    extern __shared__ __align__(16) double smem_pipe[];
    double *xs = smem_pipe;                                // [stages][ROWS][8]
    double *cs = smem_pipe + kUpdStages * ROWS * 8;        // [stages][8][kUpdStride]
    const int p8 = (p + 7) & ~7, nchunk = p8 / 8;
    const int col_tile = blockIdx.y * COLS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = WCOLS == 2 ? (warp & 3) * 32 : warp * 32, wc = WCOLS == 2 ? (warp >> 2) * 32 : 0;
    const int64_t ntiles = (n + ROWS - 1) / ROWS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row_base = tile * ROWS;
        auto issue = [&](int chunk, int stage) {
            const int k8 = chunk * 8;
            // X chunk: ROWS rows x 64 B = 4 * ROWS pieces of 16 B
#pragma unroll
            for (int it = 0; it < ROWS / 64; it++) {
                const int piece = threadIdx.x + it * 256;
                const int r = piece >> 2, part = piece & 3;
                const int64_t row = row_base + r;
                const int k = k8 + 2 * part;
                const bool ok = row < n && k < p;
                const double *src = x + (ok ? row : 0) * ldx + (ok ? k : 0);
                double *dst = xs + ((size_t)stage * ROWS + r) * 8 + 2 * part;
                if (ok && k + 1 >= p) {  // last odd column: plain stores of the single valid value
                    dst[0] = __ldg(src);
                    dst[1] = 0.0;
                } else {
                    cp_async16(dst, src, ok);
                }
            }
            // C chunk: 8 rows x COLS columns = 4 * COLS pieces of 16 B
            if (threadIdx.x < 4 * COLS) {
                const int r = threadIdx.x / (COLS / 2), part = threadIdx.x % (COLS / 2);
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
        double acc[4][NJ][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
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
            const double *xst = xs + (size_t)stage * ROWS * 8;
            const double *cst = cs + (size_t)stage * 8 * kUpdStride;
            double a0[4], a1[4], b0[NJ], b1[NJ];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double2 v = *reinterpret_cast<const double2 *>(xst + (wr + 8 * i + g) * 8 + 2 * t);
                a0[i] = v.x;
                a1[i] = v.y;
            }
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                b0[j] = cst[(2 * t) * kUpdStride + wc + 8 * j + g];
                b1[j] = cst[(2 * t + 1) * kUpdStride + wc + 8 * j + g];
            }
            // the two DMMAs of an accumulator are issued a full sweep apart (asm volatile keeps this order):
            // back to back they are a dependent pair and the warp idles for the DMMA latency after every other issue
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
        }
        cp_async_wait<0>();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = row_base + wr + 8 * i + g;
            if (r >= n) continue;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int col = col_tile + wc + 8 * j + 2 * t;
                double *yp = y + r * ldy + col;
                if (col < q) yp[0] = beta == 0.0 ? alpha * acc[i][j][0] : alpha * acc[i][j][0] + beta * yp[0];
                if (col + 1 < q) yp[1] = beta == 0.0 ? alpha * acc[i][j][1] : alpha * acc[i][j][1] + beta * yp[1];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// update, resident-C form (the wide shapes: q > 32, p <= kResMaxP).  Round-2 finding: the cp.async
// ring above synchronises the CTA once per 8-wide k-chunk; the two resident CTAs fall into lockstep
// (round-robin issue makes them finish a chunk together), so the tensor pipe idles for barrier +
// shared-memory latency once per chunk: 71 % of the DMMA peak.  Here the whole coefficient slice
// C[:, 64 columns] sits in shared memory for the CTA's lifetime, the A fragments come straight from
// global memory (64-byte row pieces, sector-exact) one chunk ahead in registers (two spill), and there is NO
// barrier inside the k loop: every warp free-runs, the scheduler interleaves warps at different
// phases.
// ---------------------------------------------------------------------------------------------
constexpr int kResMaxP = 192;

// inplace != 0: y aliases x (q <= 64, one column tile): the two warps that share a row block must both
// have read their A fragments before either stores
__global__ void __launch_bounds__(256, 2)
    update_dmma_res_kernel(int64_t n, int p, const double *x, int ldx, int q,
                           const double *__restrict__ cmat, int ldc, double alpha, double beta, double *y, int ldy,
                           int inplace) {
    extern __shared__ __align__(16) double cres[];  // (p8, kUpdStride): C[:, col tile], zero padded
    const int p8 = (p + 7) & ~7;
    const int col_tile = blockIdx.y * 64;
    for (int i = threadIdx.x; i < p8 * 64; i += blockDim.x) {
        const int k = i >> 6, cc = i & 63;
        const int col = col_tile + cc;
        cres[k * kUpdStride + cc] = (k < p && col < q) ? __ldg(cmat + (int64_t)k * ldc + col) : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int64_t ntiles = (n + 127) / 128;
    const int nfull = p / 8;  // chunks whose 8 columns all exist; a ragged last chunk is loaded with guards
    const double *cb = cres + (2 * t) * kUpdStride + wc + g;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * 128 + wr;
        const double *xr[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = r0 + 8 * i + g;
            xr[i] = x + (r < n ? r : n - 1) * ldx + 2 * t;  // rows past the end read the last row, never stored
        }
        auto load = [&](int ch, double2 (&a)[4]) {
            if (ch < nfull) {
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = __ldg(reinterpret_cast<const double2 *>(xr[i] + ch * 8));
            } else {
                const int ka = ch * 8 + 2 * t;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    a[i].x = ka < p ? __ldg(xr[i] + ch * 8) : 0.0;
                    a[i].y = ka + 1 < p ? __ldg(xr[i] + ch * 8 + 1) : 0.0;
                }
            }
        };
        const int nchunk = p8 / 8;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        // lane l also pulls the 128-byte lines of row l of the warp's 32 rows into L2 kPf chunks ahead
        // (one line = two chunks): the register prefetch below then waits for an L2 hit, not for DRAM
        constexpr int kPf = 6;
        const int64_t prow = r0 + lane;
        const double *pf = x + (prow < n ? prow : n - 1) * ldx;
        {  // ... and the head of the NEXT row tile of this CTA (its first kPf chunks)
            const int64_t nrow = prow + (int64_t)gridDim.x * 128;
            if (nrow < n) {
                const double *npf = x + nrow * ldx;
#pragma unroll
                for (int c2 = 0; c2 < kPf; c2 += 2)
                    if (c2 * 8 < p) asm volatile("prefetch.global.L2 [%0];" ::"l"(npf + c2 * 8));
            }
        }
        double2 a0[4], a1[4];
        load(0, a0);
        for (int ch = 0; ch < nchunk; ch++) {
            if (!(ch & 1) && (ch + kPf) * 8 < p) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (ch + kPf) * 8));
            if (ch + 1 < nchunk) load(ch + 1, a1);
            const double *cc = cb + (size_t)ch * 8 * kUpdStride;
            double b0[4], b1[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                b0[j] = cc[8 * j];
                b1[j] = cc[kUpdStride + 8 * j];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i].x, b0[j]);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i].y, b1[j]);
#pragma unroll
            for (int i = 0; i < 4; i++) a0[i] = a1[i];
        }
        if (inplace) __syncthreads();
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

// peak probe: DMMAs on register operands only (no memory): what the fp64 tensor pipe sustains
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *out) {
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
    }
    for (int it = 0; it < iters; it++) {
        // the two DMMAs of an accumulator are issued a full sweep apart (asm volatile keeps this order):
        // back to back they are a dependent pair and the warp idles for the DMMA latency after every other issue
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], b[i], a[j]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;  // keeps the loop alive
}

// benchmark aid (lb_dense_benchmark): 1 forces the 64-column tile for every q
static int g_update_wide_only = 0;
static int g_update_variant = 0;  // 2: the cp.async ring kernel for the wide shapes too (A/B aid)

void update_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *cmat, int ldc,
                 double alpha, double beta, double *y, int ldy) {
    const bool aligned = (ldx % 2 == 0) && (ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(cmat) & 15) == 0);
    if (aligned) {
        // column tile: 64 for wide right-hand sides, 32 / 16 / 8 (256-row tiles) for narrow ones
        const int cols = (q > 32 || g_update_wide_only) ? 64 : q > 16 ? 32 : q > 8 ? 16 : 8;
        // q = 64 a + r with 0 < r <= 32 (late iterations: 78, 92 active columns): the 64-wide tiles take
        // the first 64 a columns, a narrow tile the rest - a full tile on r columns is mostly padding
        if (cols == 64 && q > 64 && q % 64 != 0 && q % 64 <= 32 && !g_update_wide_only) {
            const int qm = q / 64 * 64;
            update_dmma(c, n, p, x, ldx, qm, cmat, ldc, alpha, beta, y, ldy);
            update_dmma(c, n, p, x, ldx, q - qm, cmat + qm, ldc, alpha, beta, y + qm, ldy);
            return;
        }
        if (cols == 64 && p <= kResMaxP && g_update_variant != 2) {
            const size_t smem_res = (size_t)((p + 7) & ~7) * kUpdStride * sizeof(double);
            static bool res_attr[64] = {};
            if (c->device >= 64 || !res_attr[c->device]) {
                LB_CUDA(cudaFuncSetAttribute(update_dmma_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((size_t)kResMaxP * kUpdStride * sizeof(double))));
                if (c->device < 64) res_attr[c->device] = true;
            }
            const int yt = cdiv(q, 64);
            const int64_t nt = (n + 127) / 128;
            dim3 grid_res((int)std::min<int64_t>(nt, std::max(1, (kSMs * 2) / yt)), yt);
            LB_LAUNCH(c, update_dmma_res_kernel, grid_res, 256, smem_res, n, p, x, ldx, q, cmat, ldc, alpha, beta, y, ldy,
                      (int)(x == y));
            return;
        }
        const int rows = cols == 64 ? 128 : 256;
        const size_t smem = (size_t)kUpdStages * (rows * 8 + 8 * kUpdStride) * sizeof(double);
        const int ytiles = cdiv(q, cols);
        const int64_t ntiles = (n + rows - 1) / rows;
        const int gx = (int)std::min<int64_t>(ntiles, std::max(1, (kSMs * 2) / ytiles));
        dim3 grid(gx, ytiles);
        // the dynamic shared memory limit is a per-device, per-instantiation setting: once per (device, tile)
#define LB_UPD(WC, NJ)                                                                                                     \
    do {                                                                                                                   \
        static bool attr_set[64] = {};                                                                                     \
        if (c->device >= 64 || !attr_set[c->device]) {                                                                     \
            LB_CUDA(cudaFuncSetAttribute(update_dmma_pipe_kernel<WC, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            if (c->device < 64) attr_set[c->device] = true;                                                                \
        }                                                                                                                  \
        LB_LAUNCH(c, (update_dmma_pipe_kernel<WC, NJ>), grid, 256, smem, n, p, x, ldx, q, cmat, ldc, alpha, beta, y, ldy);      \
    } while (0)
        if (cols == 64) LB_UPD(2, 4);
        else if (cols == 32) LB_UPD(1, 4);
        else if (cols == 16) LB_UPD(1, 2);
        else LB_UPD(1, 1);
#undef LB_UPD
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

// X(n, p) <- X C(p, p) in place.  Possible when one column tile covers all p columns (every CTA then
// reads the rows of its tile completely before it writes them) and the aligned kernels apply;
// returns false otherwise (the caller goes through a scratch block).
bool update_dmma_inplace(lb_ctx *c, int64_t n, int p, double *x, int ldx, const double *cmat, int ldc) {
    const bool aligned = (ldx % 2 == 0) && (ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(cmat) & 15) == 0);
    if (!aligned || p > 64 || g_update_wide_only) return false;
    update_dmma(c, n, p, x, ldx, p, cmat, ldc, 1.0, 0.0, x, ldx);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Gram: CTA = 4 warps = one 64 x QT output tile over a slab of rows; warp = 32 x QT/2 (NJ = QT/16
// 8-column blocks).  QT = 64 for the wide products; 32 / 16 for the narrow right-hand sides of the late
// LOBPCG iterations (a 64-wide tile ran the 16-column products at 1.5-4 TFLOP/s: DMMAs on padding).
// partial[(split * ntile + tile) * 64*QT + i*QT + j]; gram_reduce sums the slabs in fixed order.
// (Round 2: an L2 prefetch of the rows 128 ahead, the trick that helped the update, cost 15 % here.)
// ---------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(128, 4)
    gram_dmma_kernel(int64_t n, int p, const double *__restrict__ x, int ldx, int q, const double *__restrict__ y,
                     int ldy, int qtiles, int symmetric, int64_t rows_per_split, double *__restrict__ partial,
                     int ntile_total) {
    constexpr int QT = 16 * NJ;
    const int tile = blockIdx.x;
    const int pt = tile / qtiles, qt = tile % qtiles;
    if (symmetric && qt < pt) return;  // mirrored by gram_reduce (QT == 64 only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int pc = pt * 64 + (warp & 1) * 32, qc = qt * QT + (warp >> 1) * (QT / 2);
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(n, r_begin + rows_per_split);
    if (pc >= p || qc >= q) return;  // this warp's 32 x QT/2 piece is all padding (gram_reduce never reads it)
    double acc[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    bool aok[4], bok[NJ];
#pragma unroll
    for (int i = 0; i < 4; i++) aok[i] = pc + 8 * i + g < p;
#pragma unroll
    for (int j = 0; j < NJ; j++) bok[j] = qc + 8 * j + g < q;
    for (int64_t r8 = r_begin; r8 < r_end; r8 += 8) {
        const int64_t ra = r8 + 2 * t, rb = ra + 1;
        const bool va = ra < r_end, vb = rb < r_end;
        const double *xa = x + ra * ldx + pc + g, *xb = x + rb * ldx + pc + g;
        const double *ya = y + ra * ldy + qc + g, *yb = y + rb * ldy + qc + g;
        double a0[4], a1[4], b0[NJ], b1[NJ];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a0[i] = (va && aok[i]) ? __ldg(xa + 8 * i) : 0.0;
            a1[i] = (vb && aok[i]) ? __ldg(xb + 8 * i) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            b0[j] = (va && bok[j]) ? __ldg(ya + 8 * j) : 0.0;
            b1[j] = (vb && bok[j]) ? __ldg(yb + 8 * j) : 0.0;
        }
        // the two DMMAs of an accumulator are issued a full sweep apart (asm volatile keeps this order):
        // back to back they are a dependent pair and the warp idles for the DMMA latency after every other issue
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
    }
    double *out = partial + ((int64_t)blockIdx.y * ntile_total + tile) * (64 * QT);
    const int li = (warp & 1) * 32, lj = (warp >> 1) * (QT / 2);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2 *>(out + (li + 8 * i + g) * QT + lj + 8 * j + 2 * t) = v;
        }
}

__global__ void gram_reduce_kernel(int p, int q, int qtiles, int qt_width, int ntile_total, int nsplit, int symmetric,
                                   const double *__restrict__ partial, double *__restrict__ cmat) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * q) return;
    const int i = idx / q, j = idx % q;
    int pi = i, pj = j;
    if (symmetric && (j >> 6) < (i >> 6)) {  // lower tile: take the transposed entry of the upper tile
        pi = j;
        pj = i;
    }
    const int tile = (pi >> 6) * qtiles + pj / qt_width;
    const int64_t tsize = 64 * qt_width;
    const double *src = partial + (int64_t)tile * tsize + (pi & 63) * qt_width + pj % qt_width;
    double s = 0.0;
    for (int k = 0; k < nsplit; k++) s += src[(int64_t)k * ntile_total * tsize];
    cmat[idx] = s;
}

// symmetric != 0: the caller guarantees X^T Y is symmetric (S^T (A S), W^T (B W)) and p == q
void gram_dmma(lb_ctx *c, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy, double *cmat,
               bool symmetric) {
    if (p <= 64 && q <= 64) symmetric = false;  // a single 64-row tile: nothing to mirror, narrow tiles allowed
    const int qt_width = (q > 32 || symmetric) ? 64 : q > 16 ? 32 : 16;
    const int ptiles = cdiv(p, 64), qtiles = cdiv(q, qt_width);
    const int ntile = ptiles * qtiles;
    const int active = symmetric ? ptiles * (ptiles + 1) / 2 : ntile;
    int nsplit = std::max(1, (kSMs * 4) / active);
    int64_t rows_per_split = ((n + nsplit - 1) / nsplit + 7) & ~7ll;
    if (rows_per_split < 512) rows_per_split = 512;
    nsplit = (int)((n + rows_per_split - 1) / rows_per_split);
    DBuf<double> partial(c, (size_t)nsplit * ntile * 64 * qt_width);
    dim3 grid(ntile, nsplit);
    if (qt_width == 64)
        LB_LAUNCH(c, gram_dmma_kernel<4>, grid, 128, 0, n, p, x, ldx, q, y, ldy, qtiles, (int)symmetric, rows_per_split,
                  partial.p, ntile);
    else if (qt_width == 32)
        LB_LAUNCH(c, gram_dmma_kernel<2>, grid, 128, 0, n, p, x, ldx, q, y, ldy, qtiles, 0, rows_per_split, partial.p, ntile);
    else
        LB_LAUNCH(c, gram_dmma_kernel<1>, grid, 128, 0, n, p, x, ldx, q, y, ldy, qtiles, 0, rows_per_split, partial.p, ntile);
    LB_LAUNCH(c, gram_reduce_kernel, cdiv(p * q, 256), 256, 0, p, q, qtiles, qt_width, ntile, nsplit, (int)symmetric,
              partial.p, cmat);
}

// device-resident timing of the dense block products (x, y, c resident): op 0 = Gram X^T Y (p x q),
// 1 = update Y = X C, 2 = register-only DMMA peak probe (returns TFLOP/s); variant 1: the update with
// the 64-column tile for every q (A/B aid for the narrow tiles)
double dense_benchmark(lb_ctx *c, int64_t n, int p, int q, int op, int variant, int reps) {
    float ms = 0;
    if (op == 2) {
        DBuf<double> out(c, 1);
        const int iters = 4096;
        LB_LAUNCH(c, dmma_peak_kernel, kSMs * 4, 256, 0, 16, out.p);
        LB_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int i = 0; i < reps; i++) LB_LAUNCH(c, dmma_peak_kernel, kSMs * 4, 256, 0, iters, out.p);
        LB_CUDA(cudaEventRecord(c->ev1, c->stream));
        LB_CUDA(cudaEventSynchronize(c->ev1));
        LB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        // flops per launch: blocks * warps * iters * 32 DMMA * (8*8*4*2)
        const double flops = (double)kSMs * 4 * 8 * iters * 32.0 * 512.0;
        return flops / (ms / reps * 1e-3) / 1e12;
    }
    DBuf<double> x(c, (size_t)n * p), y(c, (size_t)n * q), cm(c, (size_t)p * q);
    fill_random(c, n, p, x.p, p, 7);
    fill_random(c, n, q, y.p, q, 8);
    fill_random(c, p, q, cm.p, q, 9);
    const int saved = g_update_wide_only;
    if (variant == 1) g_update_wide_only = 1;
    g_update_variant = variant == 2 ? 2 : 0;
    auto run = [&]() {
        if (op == 0) gram_dmma(c, n, p, x.p, p, q, y.p, q, cm.p, false);
        else update_dmma(c, n, p, x.p, p, q, cm.p, q, 1.0, 0.0, y.p, q);
    };
    for (int i = 0; i < 2; i++) run();
    LB_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; i++) run();
    LB_CUDA(cudaEventRecord(c->ev1, c->stream));
    LB_CUDA(cudaEventSynchronize(c->ev1));
    LB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    g_update_wide_only = saved;
    g_update_variant = 0;
    return ms / reps;
}

}  // namespace lb
