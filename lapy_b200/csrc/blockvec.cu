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

template <typename T>
struct Vec2T;
template <>
struct Vec2T<double> {
    typedef double2 V;
};
template <>
struct Vec2T<float> {
    typedef float2 V;
};

template <typename T, int G, bool VEC>
__global__ void __launch_bounds__(256) spmm_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                                   const int32_t *__restrict__ indices, const T *__restrict__ val,
                                                   const T *__restrict__ x, int ldx, T *y, int ldy, int m, int mode,
                                                   const T *b, int ldb, SpmmEpilogueT<T> epi) {
    // Row-wise form: the general fallback of the strip-staged kernel below (odd column counts, unaligned
    // operands, rows too long for the strip's shared-memory budget).  The group fetches a row's (index,
    // value) pairs with ONE coalesced load each (lane q holds entry q) and broadcasts them by warp shuffle,
    // leaving only the X gathers on the L1 pipe.  ncu (round 2): 302 warp instructions per row and three
    // dependent latencies per row (row pointers -> entries -> X rows) - what the strip kernel removes.
    typedef typename Vec2T<T>::V V2;
    constexpr int GROUPS = 256 / G;  // rows in flight per CTA
    const int grp = threadIdx.x / G, lane = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int64_t strip0 = (int64_t)blockIdx.x * kSpmmStrip;
    const int nrows = (int)(min(n, strip0 + kSpmmStrip) - strip0);
    for (int lr = grp; lr < nrows; lr += GROUPS) {
        const int64_t row = strip0 + lr;
        const int beg = __ldg(indptr + row), end = __ldg(indptr + row + 1);
        // VEC: lane owns columns c0 + 2*lane, +1 (one vector load); else c0 + lane, c0 + lane + G
        for (int c0 = 0; c0 < m; c0 += 2 * G) {
            const int ca = VEC ? c0 + 2 * lane : c0 + lane;
            const int cb = VEC ? ca + 1 : ca + G;
            const bool ha = ca < m, hb = cb < m;
            T s0 = 0, s1 = 0;
            for (int p0 = beg; p0 < end; p0 += G) {
                const int cnt = min(G, end - p0);
                int jl = 0;
                T al = 0;
                if (lane < cnt) {
                    jl = __ldg(indices + p0 + lane);
                    al = __ldg(val + p0 + lane);
                }
                int q = 0;
                for (; q + 3 < cnt; q += 4) {
                    int j[4];
                    T a[4], u0[4], u1[4];
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        j[w] = __shfl_sync(gmask, jl, q + w, G);
                        a[w] = __shfl_sync(gmask, al, q + w, G);
                    }
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const T *xr = x + (int64_t)j[w] * ldx;
                        if (VEC && hb) {
                            const V2 v = __ldg(reinterpret_cast<const V2 *>(xr + ca));
                            u0[w] = v.x;
                            u1[w] = v.y;
                        } else {
                            u0[w] = ha ? __ldg(xr + ca) : (T)0;
                            u1[w] = hb ? __ldg(xr + cb) : (T)0;
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
                    const T a0 = __shfl_sync(gmask, al, q, G);
                    const T *xr = x + (int64_t)j0 * ldx;
                    if (VEC && hb) {
                        const V2 v = __ldg(reinterpret_cast<const V2 *>(xr + ca));
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
                const T di = epi.c2 * __ldg(epi.dinv + row);
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
                const T di = epi.c2 * __ldg(epi.dinv + row);
                if (ha) {
                    const T dold = x[row * ldx + ca];
                    const T dn = fma(epi.c1, dold, di * (b[row * ldb + ca] - s0));
                    T *sp = epi.out2 + row * epi.ldout2 + ca;
                    const T v = (epi.overwrite ? (T)0 : *sp) + dold + dn;
                    if (epi.out64) epi.out64[row * epi.ldout64 + ca] = (double)v;
                    else *sp = v;
                }
                if (hb) {
                    const T dold = x[row * ldx + cb];
                    const T dn = fma(epi.c1, dold, di * (b[row * ldb + cb] - s1));
                    T *sp = epi.out2 + row * epi.ldout2 + cb;
                    const T v = (epi.overwrite ? (T)0 : *sp) + dold + dn;
                    if (epi.out64) epi.out64[row * epi.ldout64 + cb] = (double)v;
                    else *sp = v;
                }
                continue;
            }
            if (ha) y[row * ldy + ca] = s0;
            if (hb) y[row * ldy + cb] = s1;
        }
    }
}

// fallback launch for either value type
template <typename T>
static void spmm_rowwise(lb_ctx *c, const lb_mat *a, const T *v, const T *x, int ldx, T *y, int ldy, int m, int mode,
                         const T *b, int ldb, const SpmmEpilogueT<T> &epi) {
    const int32_t *ip = a->indptr.p, *ix = a->indices.p;
    const int64_t n = a->n;
    // vector loads of X need an even leading dimension and an aligned base
    const bool vec = (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & (2 * sizeof(T) - 1)) == 0);
    const int grid = cdiv(n, kSpmmStrip);
#define LB_SPMM(G)                                                                                                        \
    do {                                                                                                                  \
        if (vec) LB_LAUNCH(c, (spmm_kernel<T, G, true>), grid, 256, 0, n, ip, ix, v, x, ldx, y, ldy, m, mode, b, ldb, epi);  \
        else LB_LAUNCH(c, (spmm_kernel<T, G, false>), grid, 256, 0, n, ip, ix, v, x, ldx, y, ldy, m, mode, b, ldb, epi);     \
    } while (0)
    if (m <= 8) LB_SPMM(4);
    else if (m <= 16) LB_SPMM(8);
    else if (m <= 32) LB_SPMM(16);
    else LB_SPMM(32);
#undef LB_SPMM
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

// ---- SpMM, strip-staged form --------------------------------------------------------------------
// Round-2 finding (profiles/spmm_shapes_r2.json): the time of spmm_kernel above is rows x 0.28 us/Mrow +
// nnz x 0.029 us/Mnnz - at 7 entries per row (triangle meshes) 58 % of it is per-ROW cost: three
// dependent latencies (row pointers -> entries -> X rows) and 302 warp instructions per row.
// Here a CTA stages the CSR segment of its strip of R rows in shared memory with two coalesced
// passes (the only dependent latencies, paid once per CTA and hidden by the other resident CTAs),
// after which a row costs one shared-memory read per entry (broadcast), up to 8 independent 16-byte
// X gathers in flight, and the epilogue.  Same summation order as spmm_kernel (CSR order, one fma
// per entry) -> bit-identical results.  T = double (double2 per lane) or float (float4 per lane:
// the single-precision multigrid cycle of the eigensolver's preconditioner, amg.cu).
template <typename T>
struct VecT;
template <>
struct VecT<double> {
    typedef double2 V;
    static constexpr int W = 2;
};
template <>
struct VecT<float> {
    typedef float4 V;
    static constexpr int W = 4;
};
__device__ __forceinline__ void vfma(double2 &acc, double a, const double2 &u) {
    acc.x = fma(a, u.x, acc.x);
    acc.y = fma(a, u.y, acc.y);
}
__device__ __forceinline__ void vfma(float4 &acc, float a, const float4 &u) {
    acc.x = fmaf(a, u.x, acc.x);
    acc.y = fmaf(a, u.y, acc.y);
    acc.z = fmaf(a, u.z, acc.z);
    acc.w = fmaf(a, u.w, acc.w);
}
template <typename T, typename V>
__device__ __forceinline__ void vunpack(const V &v, T *s);
template <>
__device__ __forceinline__ void vunpack<double, double2>(const double2 &v, double *s) {
    s[0] = v.x;
    s[1] = v.y;
}
template <>
__device__ __forceinline__ void vunpack<float, float4>(const float4 &v, float *s) {
    s[0] = v.x;
    s[1] = v.y;
    s[2] = v.z;
    s[3] = v.w;
}
__device__ __forceinline__ double2 vpack(const double *s) { return make_double2(s[0], s[1]); }
__device__ __forceinline__ float4 vpack(const float *s) { return make_float4(s[0], s[1], s[2], s[3]); }

constexpr int kStripPad = 8;  // entries the 8-wide inner loop may address past the end of the strip

template <typename T, int G, int MODE, int CH, int MINB>
__global__ void __launch_bounds__(256, MINB) spmm_strip_kernel(int64_t n, int R, int cap, const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices,
                                                         const T *__restrict__ val, const T *__restrict__ x, int ldx,
                                                         T *y, int ldy, int m, const T *b, int ldb,
                                                         SpmmEpilogueT<T> epi) {
    using V = typename VecT<T>::V;
    constexpr int W = VecT<T>::W;
    constexpr int GROUPS = 256 / G;
    extern __shared__ __align__(16) unsigned char strip_smem[];
    int32_t *s_ptr = reinterpret_cast<int32_t *>(strip_smem);  // R + 1 row pointers
    T *s_val = reinterpret_cast<T *>(strip_smem + (size_t)((R + 4) & ~3) * 4);
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_val + cap + kStripPad);
    const int64_t strip0 = (int64_t)blockIdx.x * R;
    const int nrows = (int)min((int64_t)R, n - strip0);
    for (int i = threadIdx.x; i <= nrows; i += 256) s_ptr[i] = __ldg(indptr + strip0 + i);
    __syncthreads();
    const int base = s_ptr[0], total = s_ptr[nrows] - base;  // <= cap (host-side strip statistics)
    for (int i = threadIdx.x; i < total; i += 256) {
        s_idx[i] = __ldcs(indices + base + i);
        s_val[i] = __ldcs(val + base + i);
    }
    if (threadIdx.x < kStripPad) s_idx[total + threadIdx.x] = 0;
    __syncthreads();
    const int grp = threadIdx.x / G, lane = threadIdx.x % G;
    const unsigned xstride = (unsigned)ldx * (unsigned)sizeof(T);
    for (int c0 = 0; c0 < m; c0 += G * W) {
        const int col = c0 + lane * W;
        if (col >= m) continue;
        const char *xl = reinterpret_cast<const char *>(x + col);
        for (int lr = grp; lr < nrows; lr += GROUPS) {
            const int beg = s_ptr[lr] - base, end = s_ptr[lr + 1] - base;
            const int64_t row = strip0 + lr;
            V bv, dold;
            T di = 0;
            if (MODE >= 1) bv = *reinterpret_cast<const V *>(b + row * ldb + col);
            if (MODE >= 3) di = epi.c2 * __ldg(epi.dinv + row);
            if (MODE == 4) dold = *reinterpret_cast<const V *>(x + row * ldx + col);
            V acc;
            {
                T z[W] = {};
                acc = vpack(z);
            }
            for (int p = beg; p < end; p += CH) {
                // the gathers are unconditional and issued back to back: slots past the end of the row
                // address the next row's columns (valid X rows, usually the next gathers anyway; the pad
                // after the strip holds column 0), only the fmas are predicated - no predicate is live
                // across the loads, which is what let the compiler serialise them
                V u[CH];
#pragma unroll
                for (int q = 0; q < CH; q++)
                    u[q] = __ldg(reinterpret_cast<const V *>(xl + (size_t)((uint64_t)(uint32_t)s_idx[p + q] * xstride)));
                asm volatile("" ::: "memory");
#pragma unroll
                for (int q = 0; q < CH; q++)
                    if (p + q < end) vfma(acc, s_val[p + q], u[q]);
            }
            T s[W], bb[W], out[W];
            vunpack<T, V>(acc, s);
            if (MODE >= 1) vunpack<T, V>(bv, bb);
            if (MODE == 0) {
                *reinterpret_cast<V *>(y + row * ldy + col) = acc;
            } else if (MODE == 1 || MODE == 3) {  // residual (3: + first Chebyshev direction d = c2 dinv o r)
#pragma unroll
                for (int e = 0; e < W; e++) out[e] = bb[e] - s[e];
                *reinterpret_cast<V *>(y + row * ldy + col) = vpack(out);
                if (MODE == 3) {
#pragma unroll
                    for (int e = 0; e < W; e++) out[e] = di * out[e];
                    *reinterpret_cast<V *>(epi.out2 + row * epi.ldout2 + col) = vpack(out);
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int e = 0; e < W; e++) out[e] = bb[e] + s[e];
                *reinterpret_cast<V *>(y + row * ldy + col) = vpack(out);
            } else {  // 4: last Chebyshev step, x gathers d: sol (+)= d + [c1 d + c2 dinv o (b - K d)]
                T dd[W], sp[W] = {};
                vunpack<T, V>(dold, dd);
                if (!epi.overwrite) vunpack<T, V>(*reinterpret_cast<const V *>(epi.out2 + row * epi.ldout2 + col), sp);
#pragma unroll
                for (int e = 0; e < W; e++) {
                    const T dn = fma(epi.c1, dd[e], di * (bb[e] - s[e]));
                    out[e] = sp[e] + dd[e] + dn;
                }
                if (epi.out64) {  // result leaves the single-precision cycle as doubles
                    double *o = epi.out64 + row * epi.ldout64 + col;
#pragma unroll
                    for (int e = 0; e < W; e += 2) *reinterpret_cast<double2 *>(o + e) = make_double2((double)out[e], (double)out[e + 1]);
                } else {
                    *reinterpret_cast<V *>(epi.out2 + row * epi.ldout2 + col) = vpack(out);
                }
            }
        }
    }
}

// max number of stored entries in any strip of R = 8, 16, 32, 64, 128 consecutive rows
__global__ void strip_stats_kernel(int64_t n, const int32_t *__restrict__ indptr, int32_t *__restrict__ out) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (s >= n) return;
    const int beg = indptr[s];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int R = 8 << k;
        if (s % R == 0) atomicMax(out + k, indptr[min(n, s + R)] - beg);
    }
}

static void ensure_strip_stats(lb_ctx *c, const lb_mat *a) {
    if (a->strip_ready) return;
    DBuf<int32_t> d(c, 5);
    d.zero();
    LB_LAUNCH(c, strip_stats_kernel, cdiv(cdiv(a->n, 8), 256), 256, 0, a->n, a->indptr.p, d.p);
    read_back(c, a->strip_cap, d.p, 5);
    a->strip_ready = true;
}

constexpr size_t kStripSmemBudget = 32 * 1024;

// rows per CTA the strip kernel would use for this matrix and value type (0: does not fit)
template <typename T>
static int strip_rows(lb_ctx *c, const lb_mat *a, int *cap_out) {
    ensure_strip_stats(c, a);
    for (int k = 4; k >= 0; k--) {
        const int R = 8 << k;
        const size_t bytes = (size_t)((R + 4) & ~3) * 4 + (size_t)(a->strip_cap[k] + kStripPad) * (4 + sizeof(T));
        if (bytes <= kStripSmemBudget) {
            *cap_out = a->strip_cap[k];
            return R;
        }
    }
    return 0;
}

bool spmm_f32_supported(lb_ctx *, const lb_mat *a) { return !a->diagonal; }

static inline bool aligned16(const void *p, int ld, int w) {
    return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % w == 0);
}

extern thread_local int g_spmm_variant;
template <typename T, int G>
static void launch_strip(lb_ctx *c, int grid, size_t smem, int64_t n, int R, int cap, const lb_mat *a, const T *val,
                         const T *x, int ldx, T *y, int ldy, int m, int mode, const T *b, int ldb,
                         const SpmmEpilogueT<T> &epi) {
    // gathers per chunk: 4 for short rows (prolongators, ~3 entries), else 8
    const bool narrow = a->nnz <= 4 * a->n;
    // resident CTAs the kernel is compiled for (register budget vs gathers in flight), measured
    // (profiles/spmm_variants_r2.json): one row per warp (G = 32) wants 4, several rows per warp 5
    const int minb = g_spmm_variant ? g_spmm_variant : (G == 32 ? 4 : 5);
#define LB_STRIP3(MODE, CH, MINB)                                                                                   \
    LB_LAUNCH(c, (spmm_strip_kernel<T, G, MODE, CH, MINB>), grid, 256, smem, n, R, cap, a->indptr.p, a->indices.p, val, x, \
              ldx, y, ldy, m, b, ldb, epi)
#define LB_STRIP(MODE)                                                                                              \
    do {                                                                                                            \
        if (narrow) LB_STRIP3(MODE, 4, 4);                                                                          \
        else if (minb == 3) LB_STRIP3(MODE, 8, 3);                                                                  \
        else if (minb == 5) LB_STRIP3(MODE, 8, 5);                                                                  \
        else LB_STRIP3(MODE, 8, 4);                                                                                 \
    } while (0)
    switch (mode) {
        case 0: LB_STRIP(0); break;
        case 1: LB_STRIP(1); break;
        case 2: LB_STRIP(2); break;
        case 3: LB_STRIP(3); break;
        default: LB_STRIP(4); break;
    }
#undef LB_STRIP
#undef LB_STRIP3
}

// returns false when the operands do not meet the strip kernel's requirements (16-byte aligned
// rows, column count a multiple of the vector width, strip fits the shared-memory budget)
template <typename T>
static bool spmm_strip(lb_ctx *c, const lb_mat *a, const T *val, const T *x, int ldx, T *y, int ldy, int m, int mode,
                       const T *b, int ldb, const SpmmEpilogueT<T> &epi) {
    constexpr int W = VecT<T>::W;
    if (m % W || !aligned16(x, ldx, W) || !aligned16(y, ldy, W) || !aligned16(b, ldb, W) ||
        !aligned16(epi.out2, epi.ldout2, W) || !aligned16(epi.out64, epi.ldout64, 2))
        return false;
    int cap = 0;
    const int R = strip_rows<T>(c, a, &cap);
    if (R == 0) return false;
    const size_t smem = (size_t)((R + 4) & ~3) * 4 + (size_t)(cap + kStripPad) * (4 + sizeof(T));
    const int grid = cdiv(a->n, R);
    const int lanes = m / W;
    if (lanes <= 4) launch_strip<T, 4>(c, grid, smem, a->n, R, cap, a, val, x, ldx, y, ldy, m, mode, b, ldb, epi);
    else if (lanes <= 8) launch_strip<T, 8>(c, grid, smem, a->n, R, cap, a, val, x, ldx, y, ldy, m, mode, b, ldb, epi);
    else if (lanes <= 16) launch_strip<T, 16>(c, grid, smem, a->n, R, cap, a, val, x, ldx, y, ldy, m, mode, b, ldb, epi);
    else launch_strip<T, 32>(c, grid, smem, a->n, R, cap, a, val, x, ldx, y, ldy, m, mode, b, ldb, epi);
    return true;
}

// single-precision copy of the values of a matrix (made once, kept with the matrix)
__global__ void to_f32_kernel(int64_t n, const double *__restrict__ in, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

const float *mat_values_f32(lb_ctx *c, const lb_mat *a) {
    if (a->data32.n != (size_t)a->nnz || a->data32.p == nullptr) {
        a->data32.alloc(c, (size_t)a->nnz);
        if (a->nnz) LB_LAUNCH(c, to_f32_kernel, cdiv(a->nnz, 256), 256, 0, a->nnz, a->data.p, a->data32.p);
    }
    return a->data32.p;
}

void spmm_f32(lb_ctx *c, const lb_mat *a, const float *x, int ldx, float *y, int ldy, int m, int mode, const float *b,
              int ldb, const SpmmEpilogueT<float> *epi_in) {
    SpmmEpilogueT<float> epi{};
    if (epi_in) epi = *epi_in;
    const int64_t n = a->n;
    if (n == 0 || m == 0) return;
    const int64_t xrows = a->ncols < 0 ? a->n : a->ncols;
    ProfScope prof(c, PROF_SPMM, 8.0 * a->nnz + 4.0 * (n + 1) + 4.0 * m * (xrows + n * (mode ? 2 : 1)), kProfF32 + m, a->nnz);
    LB_REQUIRE(!a->diagonal, "single-precision SpMM: diagonal matrices are not supported (internal error)");
    const float *v = mat_values_f32(c, a);
    // rows too long for the strip's shared-memory budget (the near-dense coarse levels of tet meshes): row-wise kernel
    if (!spmm_strip<float>(c, a, v, x, ldx, y, ldy, m, mode, b, ldb, epi)) spmm_rowwise<float>(c, a, v, x, ldx, y, ldy, m, mode, b, ldb, epi);
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

thread_local int g_spmm_force_rowwise = 0;  // benchmark / self-test aids, per calling thread
thread_local int g_spmm_variant = 0;  // benchmark aid (lb_spmm_benchmark variant 1): time the row-wise kernel

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
    if (!g_spmm_force_rowwise && spmm_strip<double>(c, a, v, x, ldx, y, ldy, m, mode, b, ldb, epi)) return;
    spmm_rowwise<double>(c, a, v, x, ldx, y, ldy, m, mode, b, ldb, epi);
}

// ---- self-test of the SpMM kernels (lb_spmm_selftest) --------------------------------------------
__global__ void max_diff_kernel(int64_t cnt, const double *__restrict__ a, const double *__restrict__ b,
                                unsigned long long *out /* [0] max |a-b|, [1] max |b| as double bits */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const double d = fabs(a[i] - b[i]), r = fabs(b[i]);
    if (!(d == 0.0)) atomicMax(out, (unsigned long long)__double_as_longlong(d == d ? d : 1e300));
    if (r > 0.0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(r));
}
__global__ void f64_to_f32_kernel(int64_t cnt, const double *__restrict__ a, float *__restrict__ b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) b[i] = (float)a[i];
}
__global__ void f32_to_f64_kernel(int64_t cnt, const float *__restrict__ a, double *__restrict__ b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) b[i] = (double)a[i];
}
__global__ void abs_plus_one_kernel(int64_t cnt, double *a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) a[i] = 1.0 + fabs(a[i]);
}

// errs[mode] (mode 0..4): max |strip - row-wise| over all outputs of that mode in double precision (the two
// kernels sum in the same order: expected 0); errs[5 + mode]: max |single-precision strip - double| / max |double|
void spmm_selftest(lb_ctx *c, const lb_mat *a, int m, double *errs) {
    LB_REQUIRE(!a->diagonal && m % 4 == 0 && m >= 4 && (a->ncols < 0 || a->ncols == a->n), "spmm_selftest: square general matrix, m % 4 == 0");
    const int64_t n = a->n;
    const size_t cnt = (size_t)n * m;
    const int grid = cdiv(cnt, 256);
    DBuf<double> x(c, cnt), b(c, cnt), sol0(c, cnt), dinv(c, n);
    fill_random(c, n, m, x.p, m, 11);
    fill_random(c, n, m, b.p, m, 12);
    fill_random(c, n, m, sol0.p, m, 13);
    fill_random(c, n, 1, dinv.p, 1, 14);
    LB_LAUNCH(c, abs_plus_one_kernel, cdiv(n, 256), 256, 0, n, dinv.p);
    DBuf<float> xf(c, cnt), bf(c, cnt), dinvf(c, n), yf(c, cnt), o2f(c, cnt);
    LB_LAUNCH(c, f64_to_f32_kernel, grid, 256, 0, (int64_t)cnt, x.p, xf.p);
    LB_LAUNCH(c, f64_to_f32_kernel, grid, 256, 0, (int64_t)cnt, b.p, bf.p);
    LB_LAUNCH(c, f64_to_f32_kernel, cdiv(n, 256), 256, 0, n, dinv.p, dinvf.p);
    DBuf<double> y[2] = {DBuf<double>(c, cnt), DBuf<double>(c, cnt)}, o2[2] = {DBuf<double>(c, cnt), DBuf<double>(c, cnt)};
    DBuf<double> conv(c, cnt);
    DBuf<unsigned long long> mx(c, 2);
    auto diff = [&](const double *p, const double *q, double *dmax, double *rmax) {
        mx.zero();
        LB_LAUNCH(c, max_diff_kernel, grid, 256, 0, (int64_t)cnt, p, q, mx.p);
        unsigned long long h[2];
        read_back(c, h, mx.p, 2);
        double d, r;
        std::memcpy(&d, &h[0], 8);
        std::memcpy(&r, &h[1], 8);
        *dmax = std::max(*dmax, d);
        *rmax = std::max(*rmax, r);
    };
    for (int mode = 0; mode <= 4; mode++) {
        for (int v = 0; v < 2; v++) {  // 0: strip, 1: row-wise
            SpmmEpilogue e{};
            e.dinv = dinv.p;
            e.c1 = 0.375;
            e.c2 = 0.8125;
            e.out2 = o2[v].p;
            e.ldout2 = m;
            y[v].zero();
            d2d(c, o2[v].p, sol0.p, cnt * sizeof(double));
            g_spmm_force_rowwise = v;
            spmm(c, a, x.p, m, y[v].p, m, m, mode, mode ? b.p : nullptr, m, mode >= 3 ? &e : nullptr);
            g_spmm_force_rowwise = 0;
        }
        double d = 0, r = 0;
        diff(y[0].p, y[1].p, &d, &r);
        diff(o2[0].p, o2[1].p, &d, &r);
        errs[mode] = d;
        // single precision against the double row-wise result
        SpmmEpilogueT<float> ef{};
        ef.dinv = dinvf.p;
        ef.c1 = 0.375f;
        ef.c2 = 0.8125f;
        ef.out2 = o2f.p;
        ef.ldout2 = m;
        yf.zero();
        LB_LAUNCH(c, f64_to_f32_kernel, grid, 256, 0, (int64_t)cnt, sol0.p, o2f.p);
        spmm_f32(c, a, xf.p, m, yf.p, m, m, mode, mode ? bf.p : nullptr, m, mode >= 3 ? &ef : nullptr);
        double df = 0, rf = 0;
        LB_LAUNCH(c, f32_to_f64_kernel, grid, 256, 0, (int64_t)cnt, yf.p, conv.p);
        diff(conv.p, y[1].p, &df, &rf);
        LB_LAUNCH(c, f32_to_f64_kernel, grid, 256, 0, (int64_t)cnt, o2f.p, conv.p);
        diff(conv.p, o2[1].p, &df, &rf);
        errs[5 + mode] = rf > 0 ? df / rf : df;
    }
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

// residual norms without materialising the residual: out[j] = sum_i (AX[i,j] - lam[j] BX[i,j])^2 and
// out[cols + j] = sum_i BX[i,j]^2 in one pass over the two blocks (the convergence test of every LOBPCG
// iteration; the separate residual block + two dot passes moved 2.5x the bytes)
__global__ void __launch_bounds__(kDotCW *kDotRY) residual_norms_partial(int64_t n, int cols, const double *__restrict__ lam,
                                                                         const double *__restrict__ ax, int ldax,
                                                                         const double *__restrict__ bx, int ldbx,
                                                                         double *__restrict__ partial) {
    __shared__ double red[2][kDotRY][kDotCW + 1];
    const int tx = threadIdx.x % kDotCW, ty = threadIdx.x / kDotCW;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
    for (int c0 = 0; c0 < cols; c0 += kDotCW) {
        const int col = c0 + tx;
        double s = 0.0, sb = 0.0;
        if (col < cols) {
            const double l = lam[col];
            for (int64_t r = r0 + ty; r < r1; r += kDotRY) {
                const double b = bx[r * ldbx + col];
                const double res = fma(-l, b, ax[r * ldax + col]);
                s = fma(res, res, s);
                sb = fma(b, b, sb);
            }
        }
        red[0][ty][tx] = s;
        red[1][ty][tx] = sb;
        __syncthreads();
        if (ty < 2 && col < cols) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < kDotRY; k++) t += red[ty][k][tx];
            partial[(int64_t)blockIdx.x * 2 * cols + ty * cols + col] = t;
        }
        __syncthreads();
    }
}

void residual_norms(lb_ctx *c, int64_t n, int cols, const double *lam, const double *ax, int ldax, const double *bx,
                    int ldbx, double *out) {
    if (cols == 0) return;
    ProfScope prof(c, PROF_DOTS, 16.0 * n * cols);
    const int nb = (int)std::min<int64_t>(kDotBlocks, std::max<int64_t>(1, n / 64));
    DBuf<double> partial(c, (size_t)nb * 2 * cols);
    LB_LAUNCH(c, residual_norms_partial, nb, kDotCW * kDotRY, 0, n, cols, lam, ax, ldax, bx, ldbx, partial.p);
    LB_LAUNCH(c, col_dots_final, cdiv(2 * cols, 64), 64, 0, nb, 2 * cols, partial.p, out);
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

// single-precision output with the column count padded to a multiple of 4 (zero columns): the input
// block of the single-precision multigrid cycle
__global__ void residual_cols_f32_kernel(int64_t n, int ncols, int npad, const int *__restrict__ idx,
                                         const double *__restrict__ lam, const double *__restrict__ ax, int ldax,
                                         const double *__restrict__ mx, int ldmx, float *__restrict__ out, int ldout) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * npad) return;
    const int64_t row = t / npad;
    const int a = (int)(t - row * npad);
    float v = 0.f;
    if (a < ncols) {
        const int col = idx[a];
        v = (float)fma(-lam[col], mx[row * ldmx + col], ax[row * ldax + col]);
    }
    out[row * ldout + a] = v;
}

void residual_cols_f32(lb_ctx *c, int64_t n, int ncols, const int *idx, const double *lam, const double *ax, int ldax,
                       const double *mx, int ldmx, float *out, int ldout) {
    const int npad = (ncols + 3) & ~3;
    if (n * npad == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 20.0 * n * ncols);
    LB_LAUNCH(c, residual_cols_f32_kernel, cdiv(n * npad, 256), 256, 0, n, ncols, npad, idx, lam, ax, ldax, mx, ldmx, out,
              ldout);
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
