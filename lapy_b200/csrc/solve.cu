// solve.cu - block preconditioned CG behind heat.diffusion / Solver.poisson (lb_solve) and the
// public SpMM entry point.
//
// Replaces splu(hmat).solve(b0) (lapy/heat.py:226-227) and the Dirichlet elimination + splu solve
// of Solver.poisson (lapy/solver.py:848-883).  No factorisation: the operator alpha*A + beta*B is
// applied matrix-free to all right-hand sides at once (CSR SpMM); the preconditioner is Jacobi
// when the operator is strongly diagonally dominant (backward-Euler heat matrix: mass dominated)
// and one smoothed-aggregation V-cycle otherwise (Poisson).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "amg.cuh"

namespace lb {

// ---- Dirichlet handling -----------------------------------------------------------------------
// inv != NULL: the operator is stored in a locality numbering; idx are the caller's vertex ids
__global__ void mark_fixed(int64_t nfix, const int64_t *__restrict__ idx, const double *__restrict__ val,
                           const int32_t *__restrict__ inv, int *__restrict__ is_fixed, double *__restrict__ dval) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nfix) return;
    const int64_t r = inv ? inv[idx[t]] : idx[t];
    is_fixed[r] = 1;
    dval[r] = val[t];
}

__global__ void broadcast_cols(int64_t n, int m, const double *__restrict__ d, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    out[t] = d[t / m];
}

// rows / columns of fixed vertices -> identity (keeps the pattern, explicit zeros)
__global__ void mask_matrix(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                            double *__restrict__ val, const int *__restrict__ is_fixed) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int fr = is_fixed[r];
    for (int p = ptr[r]; p < ptr[r + 1]; p++) {
        const int j = idx[p];
        if (fr || is_fixed[j]) val[p] = (j == r) ? 1.0 : 0.0;
    }
}

__global__ void set_fixed_rows(int64_t n, int m, const int *__restrict__ is_fixed, const double *__restrict__ d,
                               double *__restrict__ x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t r = t / m;
    if (is_fixed[r]) x[t] = d[r];
}

// min over rows of (K_ii - sum_{j!=i} |K_ij|) / K_ii  (diagonal dominance margin)
__global__ void dominance_margin(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                 const double *__restrict__ val, unsigned long long *__restrict__ out_neg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double worst = 0.0;  // we track max of (1 - margin) >= 0
    if (i < n && ptr[i + 1] > ptr[i]) {
        double d = 0.0, off = 0.0;
        for (int p = ptr[i]; p < ptr[i + 1]; p++) {
            if (idx[p] == i) d = val[p];
            else off += fabs(val[p]);
        }
        worst = d > 0.0 ? off / d : 1e300;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0 && worst > 0.0) atomicMax(out_neg, (unsigned long long)__double_as_longlong(worst));
}

// max_i |sum_j K_ij| and max_i |K_ii|: is the constant vector (numerically) in the null space of K?
__global__ void rowsum_diag_max(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                const double *__restrict__ val, unsigned long long *__restrict__ out /* [2] */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double rs = 0.0, dg = 0.0;
    if (i < n) {
        double s = 0.0;
        for (int p = ptr[i]; p < ptr[i + 1]; p++) {
            s += val[p];
            if (idx[p] == i) dg = fabs(val[p]);
        }
        rs = fabs(s);
        if (!(rs == rs)) rs = 1e300;  // NaN entries: never treated as singular
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));
        dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (rs > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(rs));
        if (dg > 0.0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(dg));
    }
}

__global__ void add_diag_shift(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                               double *__restrict__ val, double eps) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int p = ptr[r]; p < ptr[r + 1]; p++)
        if (idx[p] == r) val[p] *= (1.0 + eps);
}

// ---- PCG scalar kernels -------------------------------------------------------------------------
__global__ void pcg_alpha(int m, const double *__restrict__ rz, const double *__restrict__ pq,
                          const int *__restrict__ active, double *__restrict__ alpha, double *__restrict__ neg_alpha) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double a = (active[j] && pq[j] > 0.0) ? rz[j] / pq[j] : 0.0;
    alpha[j] = a;
    neg_alpha[j] = -a;
}

__global__ void pcg_beta(int m, double *__restrict__ rz, const double *__restrict__ rz_new,
                         const int *__restrict__ active, double *__restrict__ beta) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    beta[j] = (active[j] && rz[j] != 0.0) ? rz_new[j] / rz[j] : 0.0;
    rz[j] = rz_new[j];
}

// ---- componentwise-accurate Jacobi for strongly diagonally dominant M-matrices (heat) ----------
// x_new_i = (b_i - sum_{j != i} K_ij x_j) / K_ii.  For K = B_lumped + t A with non-positive
// off-diagonals every term is non-negative: no cancellation, the iterates increase monotonically
// from 0 and converge with COMPONENTWISE relative accuracy - which the heat method needs
// (compute_geodesic_f normalises grad u, and u spans hundreds of orders of magnitude; a Krylov
// method only controls the global energy norm and leaves the far field as noise).  A sparse LU
// (the reference, lapy/heat.py:226) has the same componentwise accuracy.
// Kernel form: the CSR-stream SpMV of blockvec.cu (products staged in shared memory, CSR arrays
// streamed with the evict-first hint, x gathered through L1/L2) over a copy of the values whose
// diagonal entries are zeroed (voff), with the Jacobi update as the row epilogue.
constexpr int kJacCap = 2944;
constexpr int kJacBatch = 2;
constexpr int kJacDepCap = 24;  // neighbour strips tracked per strip (active set)

// voff <- off-diagonal part of K (diagonal entries set to +0.0), kdiag <- diagonal of K
__global__ void split_diagonal(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                               double *__restrict__ voff, double *__restrict__ kdiag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0.0;
    for (int p = indptr[r]; p < indptr[r + 1]; p++)
        if (indices[p] == r) {
            d = voff[p];
            voff[p] = 0.0;
        }
    kdiag[r] = d;
}

template <int MC, int ROWS>
__global__ void __launch_bounds__(256) jacobi_stream_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                                            const int32_t *__restrict__ indices,
                                                            const double *__restrict__ voff,
                                                            const double *__restrict__ kdiag,
                                                            const double *__restrict__ x, const double *__restrict__ b,
                                                            double *__restrict__ y, int ld,
                                                            unsigned long long *__restrict__ max_rel,
                                                            const unsigned char *__restrict__ age_in,
                                                            unsigned char *__restrict__ age_out,
                                                            const int32_t *__restrict__ dep,
                                                            const int32_t *__restrict__ dep_cnt, double thr) {
    __shared__ double s_prod[MC][kJacCap];
    __shared__ int32_t s_ptr[ROWS + 1];
    // Active set: a strip is swept only if one of the strips its columns point into (itself included)
    // changed by more than thr (relative) in one of the last two sweeps.  Ahead of the diffusion front
    // everything is still exactly 0, behind it the values have converged: only the band in between is
    // worked on.  age = sweeps since the strip last changed (two ping-pong buffers: after two quiet
    // sweeps both hold the same values to within thr).  dep_cnt < 0: too many neighbour strips, always swept.
    if (dep) {
        __shared__ int s_active;
        if (threadIdx.x == 0) s_active = dep_cnt[blockIdx.x] < 0;
        __syncthreads();
        // (int): with nd = -1 (too many neighbour strips, always swept) the unsigned comparison was true for every
        // thread, which then read dependency slots nobody wrote - found by memcheck once the pool handed out memory
        // with large stale words (level 9; the slots used to hold small stale values, i.e. valid strip ids)
        const int nd = dep_cnt[blockIdx.x];
        if ((int)threadIdx.x < nd && age_in[dep[(int64_t)blockIdx.x * kJacDepCap + threadIdx.x]] <= 1) s_active = 1;
        __syncthreads();
        if (!s_active) {
            if (threadIdx.x == 0) {
                const int a = age_in[blockIdx.x];
                age_out[blockIdx.x] = (unsigned char)min(a + 1, 255);
            }
            return;
        }
    }
    const int64_t strip0 = (int64_t)blockIdx.x * ROWS;
    const int nrows = (int)(min(n, strip0 + ROWS) - strip0);
    for (int i = threadIdx.x; i <= nrows; i += 256) s_ptr[i] = __ldg(indptr + strip0 + i);
    __syncthreads();
    const int base = s_ptr[0], total = s_ptr[nrows] - base;
    const bool staged = total <= kJacCap;
    if (staged) {
        for (int i0 = 0; i0 < total; i0 += 256 * kJacBatch) {
            int j[kJacBatch];
            double a[kJacBatch];
#pragma unroll
            for (int u = 0; u < kJacBatch; u++) {
                const int i = i0 + u * 256 + threadIdx.x;
                j[u] = i < total ? __ldcs(indices + base + i) : 0;
                a[u] = i < total ? __ldcs(voff + base + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kJacBatch; u++) {
                const int i = i0 + u * 256 + threadIdx.x;
                if (i < total) {
                    s_prod[0][i] = a[u] * __ldg(x + (int64_t)j[u] * ld);
                    if (MC > 1) s_prod[MC - 1][i] = a[u] * __ldg(x + (int64_t)j[u] * ld + 1);
                }
            }
        }
    }
    __syncthreads();
    double rel = 0.0;
    if (threadIdx.x < nrows) {
        const int64_t row = strip0 + threadIdx.x;
        const int beg = s_ptr[threadIdx.x], end = s_ptr[threadIdx.x + 1];
        const double diag = __ldg(kdiag + row);
#pragma unroll
        for (int col = 0; col < MC; col++) {
            double off = 0.0;
            if (staged) {
                for (int p = beg; p < end; p++) off += s_prod[col][p - base];
            } else {
                for (int p = beg; p < end; p++) off += __ldg(voff + p) * __ldg(x + (int64_t)__ldg(indices + p) * ld + col);
            }
            const double xo = x[row * ld + col];
            const double xn = diag > 0.0 ? (b[row * ld + col] - off) / diag : 0.0;
            y[row * ld + col] = xn;
            // a non-finite iterate (divergence, non-finite input) must read as "not converged":
            // fmax() drops NaN, so it is mapped to +inf explicitly
            if (!(fabs(xn) <= 1.79769313486231570e308)) rel = __longlong_as_double(0x7ff0000000000000ll);
            else if (xn != 0.0) rel = fmax(rel, fabs(xn - xo) / fabs(xn));
        }
    }
    if (dep) {
        const int changed = __syncthreads_or(rel > thr);
        if (threadIdx.x == 0) {
            const int a = age_in[blockIdx.x];
            age_out[blockIdx.x] = changed ? 0 : (unsigned char)min(a + 1, 255);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) rel = fmax(rel, __shfl_xor_sync(0xffffffffu, rel, o));
    if ((threadIdx.x & 31) == 0 && rel > 0.0) atomicMax(max_rel, (unsigned long long)__double_as_longlong(rel));
}

// neighbour strips of every strip of ROWS rows: the distinct values of column / ROWS over the
// strip's entries (bitmap in shared memory, then compaction); more than kJacDepCap -> dep_cnt = -1
template <int ROWS>
__global__ void __launch_bounds__(256) strip_deps_kernel(int64_t n, const int32_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices, int nstrips,
                                                         int32_t *__restrict__ dep, int32_t *__restrict__ dep_cnt) {
    extern __shared__ unsigned s_bits[];  // ceil(nstrips / 32) words
    __shared__ int s_cnt;
    const int nwords = (nstrips + 31) / 32;
    for (int i = threadIdx.x; i < nwords; i += 256) s_bits[i] = 0u;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t strip0 = (int64_t)blockIdx.x * ROWS;
    const int64_t rlast = min(n, strip0 + ROWS);
    const int beg = indptr[strip0], end = indptr[rlast];
    for (int p = beg + threadIdx.x; p < end; p += 256) {
        const int sj = indices[p] / ROWS;
        atomicOr(&s_bits[sj >> 5], 1u << (sj & 31));
    }
    if (threadIdx.x == 0) atomicOr(&s_bits[blockIdx.x >> 5], 1u << (blockIdx.x & 31));  // itself
    __syncthreads();
    for (int i = threadIdx.x; i < nwords; i += 256) {
        unsigned w = s_bits[i];
        while (w) {
            const int bit = __ffs(w) - 1;
            w &= w - 1;
            const int slot = atomicAdd(&s_cnt, 1);
            if (slot < kJacDepCap) dep[(int64_t)blockIdx.x * kJacDepCap + slot] = i * 32 + bit;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) dep_cnt[blockIdx.x] = s_cnt <= kJacDepCap ? s_cnt : -1;
}

// active-set state of one componentwise Jacobi solve (see jacobi_stream_kernel)
struct JacobiActive {
    int rows = 0, nstrips = 0;
    DBuf<int32_t> dep, dep_cnt;
    DBuf<unsigned char> age[2];
    int cur = 0;
    bool on = false;
};

static void jacobi_active_setup(lb_ctx *c, const lb_mat *K, JacobiActive &ja) {
    const int64_t n = K->n;
    ja.rows = K->nnz <= 10 * n ? 256 : 128;  // the strip size jacobi_sweep picks
    ja.nstrips = cdiv(n, ja.rows);
    ja.on = ja.nstrips >= 64 && ja.nstrips <= 8 * 48 * 1024;  // bitmap <= 48 KB of shared memory
    if (!ja.on) return;
    ja.dep.alloc(c, (size_t)ja.nstrips * kJacDepCap);
    ja.dep_cnt.alloc(c, ja.nstrips);
    ja.age[0].alloc(c, ja.nstrips);
    ja.age[1].alloc(c, ja.nstrips);
    const size_t smem = (size_t)((ja.nstrips + 31) / 32) * sizeof(unsigned);
    if (ja.rows == 256)
        LB_LAUNCH(c, strip_deps_kernel<256>, ja.nstrips, 256, smem, n, K->indptr.p, K->indices.p, ja.nstrips, ja.dep.p, ja.dep_cnt.p);
    else
        LB_LAUNCH(c, strip_deps_kernel<128>, ja.nstrips, 256, smem, n, K->indptr.p, K->indices.p, ja.nstrips, ja.dep.p, ja.dep_cnt.p);
}

static void jacobi_active_reset(lb_ctx *c, JacobiActive &ja) {  // new right-hand sides: everything is swept twice
    if (!ja.on) return;
    ja.age[0].zero();
    ja.age[1].zero();
    ja.cur = 0;
}

static void jacobi_sweep(lb_ctx *c, const lb_mat *K, const double *voff, const double *kdiag, const double *x,
                         const double *b, double *y, int ld, int mc, unsigned long long *max_rel, JacobiActive &ja,
                         double thr) {
    const int64_t n = K->n;
    const bool wide = K->nnz <= 10 * n;  // 256 rows per CTA fit the staging buffer on average
    const unsigned char *age_in = ja.on ? ja.age[ja.cur].p : nullptr;
    unsigned char *age_out = ja.on ? ja.age[ja.cur ^ 1].p : nullptr;
    const int32_t *dep = ja.on ? ja.dep.p : nullptr, *dep_cnt = ja.on ? ja.dep_cnt.p : nullptr;
#define LB_JAC(MC, ROWS)                                                                                              \
    LB_LAUNCH(c, (jacobi_stream_kernel<MC, ROWS>), cdiv(n, ROWS), 256, 0, n, K->indptr.p, K->indices.p, voff, kdiag, x, b, y, \
              ld, max_rel, age_in, age_out, dep, dep_cnt, thr)
    if (mc == 1 && wide) LB_JAC(1, 256);
    else if (mc == 1) LB_JAC(1, 128);
    else if (wide) LB_JAC(2, 256);
    else LB_JAC(2, 128);
#undef LB_JAC
    ja.cur ^= 1;
}

struct SolveStats {
    int iterations = 0, converged = 0, levels = 0;
    double residual = 0, setup_ms = 0, solve_ms = 0;
};

// Solves K x = rhs (all device, (n,m) row-major with ld = m). K SPD (or PSD with constant null
// space when project != 0).  x is overwritten (initial guess 0).
static SolveStats block_pcg(lb_ctx *c, lb_mat *K, const double *rhs, double *x, int m, double tol, int maxit,
                            bool project, int force_prec, bool try_jacobi) {
    const int64_t n = K->n;
    SolveStats st;
    // ---- preconditioner choice
    DBuf<unsigned long long> dom(c, 1);
    dom.zero();
    LB_LAUNCH(c, dominance_margin, cdiv(n, 256), 256, 0, n, K->indptr.p, K->indices.p, K->data.p, dom.p);
    unsigned long long bits = 0;
    read_back(c, &bits, dom.p, 1);
    double off_ratio;
    std::memcpy(&off_ratio, &bits, 8);
    bool use_amg = !(off_ratio < 0.95);  // kappa(D^-1 K) <= (1+r)/(1-r) < 39
    if (c->trace) fprintf(stderr, "[lb trace] solve: n=%lld m=%d off-diagonal ratio %.4f project=%d\n", (long long)n, m, off_ratio, (int)project);
    if (force_prec == 1) use_amg = false;
    if (force_prec == 2) use_amg = true;
    // (not gated on off_ratio < 1: float32 tet meshes with obtuse elements have rows with
    // off_ratio >= 1 and still contract - measured on data/cubeTetra.vtk; a non-contracting or
    // non-finite iteration is detected below and falls through to PCG)
    if (force_prec == 0 && !project && try_jacobi) {
        // componentwise Jacobi (see jacobi_stream_kernel), two columns at a time; checks the max
        // relative increment every 64 sweeps; falls through to PCG if it does not contract
        // (input that is not an M-matrix)
        cudaEvent_t j0, j1;
        LB_CUDA(cudaEventCreate(&j0));
        LB_CUDA(cudaEventCreate(&j1));
        LB_CUDA(cudaEventRecord(j0, c->stream));
        DBuf<double> xa(c, (size_t)n * m), xb(c, (size_t)n * m);
        DBuf<unsigned long long> mr(c, 1);
        mr.zero();  // the unchecked sweeps atomicMax into it too (initcheck)
        DBuf<double> voff(c, (size_t)K->nnz), kdiag(c, n);
        d2d(c, voff.p, K->data.p, (size_t)K->nnz * sizeof(double));
        LB_LAUNCH(c, split_diagonal, cdiv(n, 256), 256, 0, n, K->indptr.p, K->indices.p, voff.p, kdiag.p);
        const double jtol = std::max(0.1 * tol, 1e-13);  // on the relative increment per sweep
        JacobiActive ja;
        jacobi_active_setup(c, K, ja);
        bool ok = true;
        int sweeps_max = 0;
        double rel_max = 0.0;
        for (int c0 = 0; c0 < m && ok; c0 += 2) {
            const int mc = std::min(2, m - c0);
            xa.zero();
            xb.zero();  // strips that are never swept (ahead of the front) must read 0 in both buffers
            jacobi_active_reset(c, ja);
            double *cur = xa.p + c0, *nxt = xb.p + c0;
            const double *bb = rhs + c0;
            double rel = 1.0, prev = 2.0;
            int sweeps = 0, stalls = 0, since_best = 0, growing = 0;
            double best = INFINITY;
            ok = false;
            while (sweeps < 100000) {
                for (int i = 0; i < 63; i++) {
                    jacobi_sweep(c, K, voff.p, kdiag.p, cur, bb, nxt, m, mc, mr.p, ja, jtol);
                    std::swap(cur, nxt);
                }
                mr.zero();
                jacobi_sweep(c, K, voff.p, kdiag.p, cur, bb, nxt, m, mc, mr.p, ja, jtol);
                std::swap(cur, nxt);
                sweeps += 64;
                unsigned long long bits2 = 0;
                read_back(c, &bits2, mr.p, 1);
                std::memcpy(&rel, &bits2, 8);
                if (c->trace && (sweeps % 512 == 0 || rel <= jtol))
                    fprintf(stderr, "[lb trace] jacobi cols %d..: %d sweeps, max relative increment %.3e (target %.1e)\n", c0,
                            sweeps, rel, jtol);
                // converged: increment below the target, or sitting on the rounding floor (operators
                // with positive off-diagonals - obtuse elements - have cancellation in b - N x): below
                // 1e-9 the smallest increment seen has not improved by 10 % over four checks.  A slowly
                // but steadily contracting increment (large t: 0.6-0.95 per 64 sweeps) keeps iterating.
                if (rel <= jtol) {
                    ok = true;
                    break;
                }
                if (!std::isfinite(rel)) break;  // overflow / NaN: fall through to PCG
                // rel >= 1 is normal while the front still reaches new vertices (at most ~diameter
                // sweeps), but it must not keep GROWING: a divergent iteration is cut after 16 checks
                if (rel >= 1.0 && rel > prev && ++growing > 16) break;
                if (rel < 1e-9) {
                    if (rel < 0.9 * best) {
                        best = rel;
                        since_best = 0;
                    } else if (++since_best >= 4) {
                        ok = true;
                        break;
                    }
                } else if (rel < 1.0 && rel >= prev && ++stalls > 8) {
                    break;  // rel stays >= 1 while the front still reaches new vertices; afterwards it must contract
                }
                prev = rel;
            }
            if (ok) copy_cols(c, n, mc, cur, m, x + c0, m);
            sweeps_max = std::max(sweeps_max, sweeps);
            rel_max = std::max(rel_max, rel);
        }
        LB_CUDA(cudaEventRecord(j1, c->stream));
        LB_CUDA(cudaEventSynchronize(j1));
        float jms = 0;
        cudaEventElapsedTime(&jms, j0, j1);
        cudaEventDestroy(j0);
        cudaEventDestroy(j1);
        if (ok) {
            st.iterations = sweeps_max;
            st.converged = m;
            st.residual = rel_max;
            st.solve_ms = jms;
            st.levels = -1;  // marks the componentwise Jacobi path
            return st;
        }
    }
    std::unique_ptr<Amg> amg;
    DBuf<double> dinv;
    if (use_amg) {
        auto Kp = mat_axpby(c, K, 1.0, nullptr, 0.0);
        if (project) LB_LAUNCH(c, add_diag_shift, cdiv(n, 256), 256, 0, n, Kp->indptr.p, Kp->indices.p, Kp->data.p, 1e-6);
        AmgOptions opt;
        amg = amg_setup(c, std::move(Kp), m, opt);
        st.levels = (int)amg->levels.size();
        st.setup_ms = amg->setup_ms;
    } else {
        DBuf<double> diag(c, n);
        extract_diagonal(c, K, diag.p);
        dinv.alloc(c, n);
        launch_diag_inverse(c, n, diag.p, dinv.p);
    }
    auto precond = [&](const double *r, double *z) {
        if (use_amg) amg_apply(*amg, r, m, z, m, m);
        else scale_rows(c, n, m, dinv.p, r, m, z, m);
        if (project) remove_col_means(c, n, m, z, m);
    };

    cudaEvent_t e0, e1;
    LB_CUDA(cudaEventCreate(&e0));
    LB_CUDA(cudaEventCreate(&e1));
    LB_CUDA(cudaEventRecord(e0, c->stream));

    DBuf<double> r(c, (size_t)n * m), z(c, (size_t)n * m), p(c, (size_t)n * m), q(c, (size_t)n * m);
    DBuf<double> rz(c, m), rz_new(c, m), pq(c, m), rr(c, m), alpha(c, m), nalpha(c, m), beta(c, m);
    DBuf<int> active(c, m);
    std::vector<double> h_rr(m), h_bb(m);
    std::vector<int> h_active(m, 1);

    d2d(c, r.p, rhs, (size_t)n * m * sizeof(double));
    if (project) remove_col_means(c, n, m, r.p, m);
    LB_CUDA(cudaMemsetAsync(x, 0, (size_t)n * m * sizeof(double), c->stream));
    col_dots(c, n, m, r.p, m, r.p, m, rr.p);
    read_back(c, h_bb.data(), rr.p, m);
    int nact = 0;
    for (int j = 0; j < m; j++) {
        h_active[j] = h_bb[j] > 0.0;
        nact += h_active[j];
    }
    h2d(c, active.p, h_active.data(), m * sizeof(int));
    sync(c);
    double worst = 0.0;
    if (nact) {
        precond(r.p, z.p);
        d2d(c, p.p, z.p, (size_t)n * m * sizeof(double));
        col_dots(c, n, m, r.p, m, z.p, m, rz.p);
        for (int it = 0; it < maxit; it++) {
            st.iterations = it + 1;
            spmm(c, K, p.p, m, q.p, m, m);
            // singular case: iterate with P K P (P = I - 11^T/n) so that residuals stay in range(K)
            // even when K 1 != 0 by rounding (fp32-assembled operators, lambda_0 ~ -1e-6)
            if (project) remove_col_means(c, n, m, q.p, m);
            col_dots(c, n, m, p.p, m, q.p, m, pq.p);
            LB_LAUNCH(c, pcg_alpha, cdiv(m, 64), 64, 0, m, rz.p, pq.p, active.p, alpha.p, nalpha.p);
            axpby_cols(c, n, m, alpha.p, 0.0, p.p, m, nullptr, 1.0, x, m);
            axpby_cols(c, n, m, nalpha.p, 0.0, q.p, m, nullptr, 1.0, r.p, m);
            col_dots(c, n, m, r.p, m, r.p, m, rr.p);
            read_back(c, h_rr.data(), rr.p, m);
            worst = 0.0;
            nact = 0;
            for (int j = 0; j < m; j++) {
                if (h_bb[j] <= 0.0) continue;
                const double rel = std::sqrt(h_rr[j] / h_bb[j]);
                if (!(rel <= tol)) {
                    nact++;
                    worst = std::max(worst, rel);
                    if (!std::isfinite(rel)) worst = INFINITY;
                } else {
                    h_active[j] = 0;
                }
            }
            if (nact == 0 || !std::isfinite(worst)) break;
            h2d(c, active.p, h_active.data(), m * sizeof(int));
            precond(r.p, z.p);
            col_dots(c, n, m, r.p, m, z.p, m, rz_new.p);
            LB_LAUNCH(c, pcg_beta, cdiv(m, 64), 64, 0, m, rz.p, rz_new.p, active.p, beta.p);
            axpby_cols(c, n, m, nullptr, 1.0, z.p, m, beta.p, 0.0, p.p, m);
        }
    }
    if (project) remove_col_means(c, n, m, x, m);
    LB_CUDA(cudaEventRecord(e1, c->stream));
    LB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    st.solve_ms = ms;
    st.residual = worst;
    st.converged = m - nact;
    return st;
}

// mean length of the unique edges = strict upper triangle of the stiffness pattern
// (TriaMesh.avg_edge_length lapy/tria_mesh.py:735-748: triu(adj_sym, 1)); edge vectors and lengths
// in fp64, fixed-tree reduction
__global__ void edge_length_partial(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                    const D4 *__restrict__ v4, double *__restrict__ psum,
                                    unsigned long long *__restrict__ pcnt) {
    double s = 0.0;
    unsigned long long cnt = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const D4 a = ldg_d4(v4 + r);
        for (int p = ptr[r]; p < ptr[r + 1]; p++) {
            const int j = idx[p];
            if (j <= r) continue;
            const D4 q = ldg_d4(v4 + j);
            const double dx = a.x - q.x, dy = a.y - q.y, dz = a.z - q.z;
            s += sqrt((dx * dx + dy * dy) + dz * dz);
            cnt++;
        }
    }
    __shared__ double ss[256];
    __shared__ unsigned long long sc[256];
    ss[threadIdx.x] = s;
    sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            ss[threadIdx.x] += ss[threadIdx.x + o];
            sc[threadIdx.x] += sc[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        psum[blockIdx.x] = ss[0];
        pcnt[blockIdx.x] = sc[0];
    }
}

static double avg_edge_length_device(lb_ctx *c, lb_mesh *mesh, lb_mat *pattern) {
    const int nb = 1024;
    DBuf<double> psum(c, nb);
    DBuf<unsigned long long> pcnt(c, nb);
    // the pattern of an assembled matrix is in the mesh's locality numbering: use the vertices of that layout
    const D4 *v4 = pattern->permuted ? mesh->v4m.p : mesh->v4s->p;
    LB_LAUNCH(c, edge_length_partial, nb, 256, 0, pattern->n, pattern->indptr.p, pattern->indices.p, v4, psum.p, pcnt.p);
    std::vector<double> hs(nb);
    std::vector<unsigned long long> hc(nb);
    read_back(c, hs.data(), psum.p, nb);
    read_back(c, hc.data(), pcnt.p, nb);
    double s = 0;
    unsigned long long cnt = 0;
    for (int i = 0; i < nb; i++) {
        s += hs[i];
        cnt += hc[i];
    }
    return cnt ? s / (double)cnt : 0.0;
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_spmm(lb_ctx *c, lb_mat *mat, const double *x, int64_t m, double *y) {
    LB_API_BEGIN
    LB_REQUIRE(c && mat && x && y, "lb_spmm: NULL argument");
    LB_REQUIRE(m >= 1 && m <= 4096, "lb_spmm: bad column count");
    DeviceGuard g(c->device);
    const int64_t n = mat->n;
    DBuf<double> dx(c, (size_t)n * m), dy(c, (size_t)n * m);
    h2d(c, dx.p, x, (size_t)n * m * sizeof(double));
    if (mat->permuted) {  // stored in the mesh's locality numbering: permute x in, y out
        DBuf<double> px(c, (size_t)n * m);
        gather_rows(c, n, (int)m, mat->ord->order.p, dx.p, (int)m, px.p, (int)m);
        spmm(c, mat, px.p, (int)m, dx.p, (int)m, (int)m);
        gather_rows(c, n, (int)m, mat->ord->inv.p, dx.p, (int)m, dy.p, (int)m);
    } else {
        spmm(c, mat, dx.p, (int)m, dy.p, (int)m, (int)m);
    }
    d2h(c, y, dy.p, (size_t)n * m * sizeof(double));
    sync(c);
    LB_API_END
}

// device-resident timing of y = M x (m columns): ms per launch over `reps` launches (CUDA events)
int lb_spmm_benchmark(lb_ctx *c, lb_mat *mat0, int64_t m, int reps, int renumber, double *ms_per_launch) {
    LB_API_BEGIN
    LB_REQUIRE(c && mat0 && ms_per_launch && m >= 1 && m <= 1024 && reps >= 1, "lb_spmm_benchmark: bad argument");
    DeviceGuard g(c->device);
    const int64_t n = mat0->n;
    // renumber != 0: the numbering the solvers work in = how assembled matrices are stored;
    // renumber == 0: the same operator in the caller's numbering (diagnostic: what the locality
    // numbering buys)
    std::unique_ptr<lb_mat> perm;
    lb_mat *mat = mat0;
    if (!(renumber & 1) && mat0->permuted) {
        perm = to_caller_order(c, mat0);
        mat = perm.get();
    }
    // bits 8.. of `renumber`: kernel variant (development aid): 1 = row-wise kernel, 2 = single precision,
    // 3 / 4 / 5 = strip kernel compiled for that many resident CTAs (23 / 24 / 25: in single precision)
    const int variant = renumber >> 8;
    const bool f32 = variant == 2 || variant >= 20;
    g_spmm_force_rowwise = variant == 1;
    g_spmm_variant = variant >= 20 ? variant - 20 : (variant >= 3 ? variant : 0);
    DBuf<double> dx(c, (size_t)n * m), dy(c, (size_t)n * m);
    fill_random(c, n, (int)m, dx.p, (int)m, 42);
    DBuf<float> fx, fy;
    const int m4 = ((int)m + 3) & ~3;
    if (f32) {
        LB_REQUIRE(spmm_f32_supported(c, mat), "matrix not supported by the single-precision SpMM");
        fx.alloc(c, (size_t)n * m4);
        fy.alloc(c, (size_t)n * m4);
        fx.zero();
    }
    auto once = [&]() {
        if (f32) spmm_f32(c, mat, fx.p, m4, fy.p, m4, m4);
        else spmm(c, mat, dx.p, (int)m, dy.p, (int)m, (int)m);
    };
    for (int i = 0; i < 3; i++) once();
    LB_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; i++) once();
    LB_CUDA(cudaEventRecord(c->ev1, c->stream));
    LB_CUDA(cudaEventSynchronize(c->ev1));
    g_spmm_force_rowwise = 0;
    g_spmm_variant = 0;
    float ms = 0;
    LB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / reps;
    LB_API_END
}

// errs (10): the strip-staged SpMM against the row-wise kernel for the five epilogue modes (double:
// expected 0.0, same summation order) and the single-precision strip kernel against double (relative)
int lb_spmm_selftest(lb_ctx *c, lb_mat *mat, int64_t m, double *errs) {
    LB_API_BEGIN
    LB_REQUIRE(c && mat && errs && m >= 4 && m <= 1024, "lb_spmm_selftest: bad argument");
    DeviceGuard g(c->device);
    spmm_selftest(c, mat, (int)m, errs);
    LB_API_END
}

int lb_dense_benchmark(lb_ctx *c, int64_t n, int64_t p, int64_t q, int op, int variant, int reps, double *result) {
    LB_API_BEGIN
    LB_REQUIRE(c && result && n > 0 && p > 0 && q > 0 && p <= 384 && q <= 4096 && reps >= 1 && op >= 0 && op <= 2,
               "lb_dense_benchmark: bad argument");
    DeviceGuard g(c->device);
    *result = dense_benchmark(c, n, (int)p, (int)q, op, variant, reps);
    LB_API_END
}

int lb_avg_edge_length(lb_ctx *c, lb_mesh *mesh, lb_mat *pattern, double *out) {
    LB_API_BEGIN
    LB_REQUIRE(c && mesh && pattern && out, "lb_avg_edge_length: NULL argument");
    LB_REQUIRE(pattern->n <= mesh->nv && (!pattern->permuted || pattern->ord == mesh->ord),
               "matrix does not belong to this mesh");
    DeviceGuard g(c->device);
    *out = avg_edge_length_device(c, mesh, pattern);
    LB_API_END
}

int lb_block_gram(lb_ctx *c, int64_t n, int64_t p, const double *x, int64_t q, const double *y, double *cmat) {
    LB_API_BEGIN
    LB_REQUIRE(c && x && y && cmat && n > 0 && p > 0 && q > 0 && p <= 4096 && q <= 4096, "lb_block_gram: bad argument");
    DeviceGuard g(c->device);
    DBuf<double> dx(c, (size_t)n * p), dy(c, (size_t)n * q), dc(c, (size_t)p * q);
    h2d(c, dx.p, x, (size_t)n * p * sizeof(double));
    h2d(c, dy.p, y, (size_t)n * q * sizeof(double));
    gram(c, n, (int)p, dx.p, (int)p, (int)q, dy.p, (int)q, dc.p, false);
    d2h(c, cmat, dc.p, (size_t)p * q * sizeof(double));
    sync(c);
    LB_API_END
}

int lb_block_update(lb_ctx *c, int64_t n, int64_t p, const double *x, int64_t q, const double *cmat, double alpha,
                    double beta, double *y) {
    LB_API_BEGIN
    LB_REQUIRE(c && x && y && cmat && n > 0 && p > 0 && q > 0 && p <= 384 && q <= 4096, "lb_block_update: bad argument");
    DeviceGuard g(c->device);
    DBuf<double> dx(c, (size_t)n * p), dy(c, (size_t)n * q), dc(c, (size_t)p * q);
    h2d(c, dx.p, x, (size_t)n * p * sizeof(double));
    h2d(c, dy.p, y, (size_t)n * q * sizeof(double));
    h2d(c, dc.p, cmat, (size_t)p * q * sizeof(double));
    update(c, n, (int)p, dx.p, (int)p, (int)q, dc.p, (int)q, alpha, beta, dy.p, (int)q);
    d2h(c, y, dy.p, (size_t)n * q * sizeof(double));
    sync(c);
    LB_API_END
}

int lb_solve(lb_ctx *c, lb_mat *a, double alpha, lb_mat *b, double beta, const double *rhs, int64_t m,
             const int64_t *fix_idx, int64_t nfix, const double *fix_val, double tol, int maxit, int project_nullspace,
             double *x, lb_info *info) {
    LB_API_BEGIN
    LB_REQUIRE(c && a && rhs && x, "lb_solve: NULL argument");
    LB_REQUIRE(m >= 1 && m <= 1024, "lb_solve: number of right-hand sides must be in [1, 1024]");
    LB_REQUIRE(nfix == 0 || (fix_idx && fix_val), "lb_solve: Dirichlet arrays missing");
    DeviceGuard g(c->device);
    const int64_t n = a->n;
    if (tol <= 0) tol = 1e-12;
    if (maxit <= 0) maxit = 2000;
    const int mm = (int)m;
    for (int64_t i = 0; i < nfix; i++)
        LB_REQUIRE(fix_idx[i] >= 0 && fix_idx[i] < n, "Dirichlet index %lld out of range", (long long)fix_idx[i]);
    // everything runs in the locality numbering of the assembled operator (Morton order of its mesh:
    // the x gathers of the sweeps / SpMVs hit L1/L2); right-hand sides and Dirichlet indices are
    // mapped in, the solution is mapped back to the caller's order
    MatView va, vb;
    std::shared_ptr<lb_order> ord = common_numbering(c, a, beta != 0.0 ? b : nullptr, va, vb);
    const int32_t *to_new = ord ? ord->inv.p : nullptr;
    auto K = mat_axpby(c, va.m, alpha, beta != 0.0 ? vb.m : nullptr, beta);
    DBuf<double> d_rhs(c, (size_t)n * mm), d_x(c, (size_t)n * mm), d_sol(c, (size_t)n * mm);
    if (ord) {
        h2d(c, d_sol.p, rhs, (size_t)n * mm * sizeof(double));
        gather_rows(c, n, mm, ord->order.p, d_sol.p, mm, d_rhs.p, mm);
    } else {
        h2d(c, d_rhs.p, rhs, (size_t)n * mm * sizeof(double));
    }
    DBuf<int> is_fixed;
    DBuf<double> dval;
    if (nfix > 0) {
        is_fixed.alloc(c, n);
        is_fixed.zero();
        dval.alloc(c, n);
        dval.zero();
        DBuf<int64_t> d_idx(c, nfix);
        DBuf<double> d_val(c, nfix);
        h2d(c, d_idx.p, fix_idx, nfix * sizeof(int64_t));
        h2d(c, d_val.p, fix_val, nfix * sizeof(double));
        LB_LAUNCH(c, mark_fixed, cdiv(nfix, 256), 256, 0, nfix, d_idx.p, d_val.p, to_new, is_fixed.p, dval.p);
        // rhs <- rhs - K d  (solver.py:846), then eliminate: identity rows/cols, rhs_fixed = d
        DBuf<double> dblock(c, (size_t)n * mm);
        LB_LAUNCH(c, broadcast_cols, cdiv(n * mm, 256), 256, 0, n, mm, dval.p, dblock.p);
        spmm(c, K.get(), dblock.p, mm, d_rhs.p, mm, mm, 1, d_rhs.p, mm);
        LB_LAUNCH(c, mask_matrix, cdiv(n, 256), 256, 0, n, K->indptr.p, K->indices.p, K->data.p, is_fixed.p);
        LB_LAUNCH(c, set_fixed_rows, cdiv(n * mm, 256), 256, 0, n, mm, is_fixed.p, dval.p, d_rhs.p);
    }
    // project out the constants only if they ARE (numerically) in the null space: a user-assigned
    // nonsingular operator (A + c*B, screened Poisson) must be solved as it is, like splu does.
    // float32-assembled stiffness matrices have |A 1| ~ 1e-7 * diag: the threshold keeps them singular.
    bool project = project_nullspace != 0 && nfix == 0;
    if (project) {
        DBuf<unsigned long long> rsd(c, 2);
        rsd.zero();
        LB_LAUNCH(c, rowsum_diag_max, cdiv(n, 256), 256, 0, n, K->indptr.p, K->indices.p, K->data.p, rsd.p);
        unsigned long long hb[2];
        read_back(c, hb, rsd.p, 2);
        double rs, dg;
        std::memcpy(&rs, &hb[0], 8);
        std::memcpy(&dg, &hb[1], 8);
        project = rs <= 1e-5 * dg;
        if (c->trace) fprintf(stderr, "[lb trace] solve: max |K 1| = %.3e, max diag = %.3e -> project = %d\n", rs, dg, (int)project);
    }
    // mass-dominated operators (backward Euler heat step: diagonal B, beta != 0): componentwise Jacobi.
    // Its contraction factor is ~ 1 - (mass share of the diagonal): the heat step t = m h^2 has a share
    // of 0.2 (m = 1) ... 0.015 (m = 16), a mean-curvature-flow step (lapy/diffgeo.py:590) 1e-7 - there
    // Jacobi would need millions of sweeps and the AMG-preconditioned CG below is used.
    bool try_jacobi = beta != 0.0 && b != nullptr && b->diagonal && nfix == 0;
    if (try_jacobi) {
        DBuf<double> dk(c, n), db(c, n), sums(c, 2);
        extract_diagonal(c, K.get(), dk.p);
        extract_diagonal(c, vb.m, db.p);
        col_dots(c, n, 1, dk.p, 1, nullptr, 0, sums.p);
        col_dots(c, n, 1, db.p, 1, nullptr, 0, sums.p + 1);
        double hs[2];
        read_back(c, hs, sums.p, 2);
        try_jacobi = std::fabs(beta) * hs[1] >= 1e-3 * hs[0];
        if (c->trace) fprintf(stderr, "[lb trace] solve: mass share of the diagonal %.3e -> %s\n", std::fabs(beta) * hs[1] / hs[0], try_jacobi ? "componentwise Jacobi" : "AMG-PCG");
    }
    SolveStats st = block_pcg(c, K.get(), d_rhs.p, d_sol.p, mm, tol, maxit, project, 0, try_jacobi);
    if (nfix > 0) LB_LAUNCH(c, set_fixed_rows, cdiv(n * mm, 256), 256, 0, n, mm, is_fixed.p, dval.p, d_sol.p);
    if (ord) gather_rows(c, n, mm, ord->inv.p, d_sol.p, mm, d_x.p, mm);
    else d2d(c, d_x.p, d_sol.p, (size_t)n * mm * sizeof(double));
    d2h(c, x, d_x.p, (size_t)n * mm * sizeof(double));
    sync(c);
    if (info) {
        info->iterations = st.iterations;
        info->converged = st.converged;
        info->amg_levels = st.levels;
        info->reserved = 0;
        info->residual = st.residual;
        info->setup_ms = st.setup_ms;
        info->solve_ms = st.solve_ms;
    }
    if (st.converged < mm) {
        set_error("PCG did not converge: %d of %d right-hand sides, max relative residual %.3e after %d iterations",
                  st.converged, mm, st.residual, st.iterations);
        return LB_ERR_NOCONV;
    }
    LB_API_END
}

}  // extern "C"
