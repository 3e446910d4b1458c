// amg.cu - on-device smoothed-aggregation AMG: setup (MIS-2 aggregation, smoothed prolongator,
// Galerkin RAP through a deterministic row-wise SpGEMM) and the block V-cycle.
// See amg.cuh for what it replaces in the reference.
#include "amg.cuh"

namespace lb {

// =============================================================================================
// generic sparse helpers
// =============================================================================================
static std::unique_ptr<lb_mat> make_mat(lb_ctx *c, int64_t n, int64_t ncols, int64_t nnz) {
    auto m = std::make_unique<lb_mat>();
    m->ctx = c;
    m->n = n;
    m->ncols = ncols;
    m->nnz = nnz;
    m->indptr.alloc(c, n + 1);
    m->indices.alloc(c, nnz);
    m->data.alloc(c, nnz);
    return m;
}

__global__ void axpby_same_pattern(int64_t nnz, double alpha, const double *__restrict__ a, double beta,
                                   const double *__restrict__ b, double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) out[p] = alpha * a[p] + beta * b[p];
}

__global__ void pattern_differs(int64_t nnz, const int32_t *__restrict__ a, const int32_t *__restrict__ b, int *flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz && a[p] != b[p]) *flag = 1;
}

__global__ void axpby_diag(int64_t n, double alpha, const int32_t *__restrict__ indptr,
                           const int32_t *__restrict__ indices, const double *__restrict__ a, double beta,
                           const int32_t *__restrict__ bptr, const double *__restrict__ bval,
                           double *__restrict__ out, int *missing_diag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double bd = bptr[r + 1] > bptr[r] ? bval[bptr[r]] : 0.0;
    bool found = false;
    for (int p = indptr[r]; p < indptr[r + 1]; p++) {
        double v = alpha * a[p];
        if (indices[p] == r) {
            v += beta * bd;
            found = true;
        }
        out[p] = v;
    }
    if (!found && bd != 0.0) *missing_diag = 1;
}

std::unique_ptr<lb_mat> mat_axpby(lb_ctx *c, const lb_mat *a, double alpha, const lb_mat *b, double beta) {
    LB_REQUIRE(b == nullptr || a->n == b->n, "matrix dimensions differ (%lld vs %lld)", (long long)a->n,
               (long long)(b ? b->n : 0));
    auto out = make_mat(c, a->n, -1, a->nnz);
    d2d(c, out->indptr.p, a->indptr.p, (a->n + 1) * sizeof(int32_t));
    d2d(c, out->indices.p, a->indices.p, a->nnz * sizeof(int32_t));
    DBuf<int> flag(c, 1);
    flag.zero();
    if (b == nullptr || beta == 0.0) {
        LB_LAUNCH(c, axpby_same_pattern, cdiv(a->nnz, 256), 256, 0, a->nnz, alpha, a->data.p, 0.0, a->data.p,
                  out->data.p);
        return out;
    }
    if (b->diagonal) {
        LB_LAUNCH(c, axpby_diag, cdiv(a->n, 256), 256, 0, a->n, alpha, a->indptr.p, a->indices.p, a->data.p, beta,
                  b->indptr.p, b->data.p, out->data.p, flag.p);
    } else {
        LB_REQUIRE(a->nnz == b->nnz, "stiffness and mass have different sparsity patterns (nnz %lld vs %lld)",
                   (long long)a->nnz, (long long)b->nnz);
        LB_LAUNCH(c, pattern_differs, cdiv(a->nnz, 256), 256, 0, a->nnz, a->indices.p, b->indices.p, flag.p);
        LB_LAUNCH(c, axpby_same_pattern, cdiv(a->nnz, 256), 256, 0, a->nnz, alpha, a->data.p, beta, b->data.p,
                  out->data.p);
    }
    int h = 0;
    read_back(c, &h, flag.p, 1);
    if (h) {
        set_error("stiffness and mass matrices must share their sparsity pattern (or the mass must be diagonal)");
        throw Error{LB_ERR_UNSUPPORTED};
    }
    return out;
}

// ---- symmetric permutation P A P^T (locality renumbering; column order inside rows is kept) -----
__global__ void perm_row_len(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ order,
                             int32_t *__restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        const int o = order[r];
        len[r] = ptr[o + 1] - ptr[o];
    }
}

__global__ void perm_copy_rows(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                               const double *__restrict__ val, const int32_t *__restrict__ order,
                               const int32_t *__restrict__ inv, const int32_t *__restrict__ nptr,
                               int32_t *__restrict__ nidx, double *__restrict__ nval) {
    // 8 lanes per row
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int lane = threadIdx.x & 7;
    if (r >= n) return;
    const int o = order[r];
    const int src = ptr[o], len = ptr[o + 1] - src, dst = nptr[r];
    for (int q = lane; q < len; q += 8) {
        nidx[dst + q] = inv[idx[src + q]];
        nval[dst + q] = val[src + q];
    }
}

std::unique_ptr<lb_mat> permute_symmetric(lb_ctx *c, const lb_mat *a, const int32_t *order, const int32_t *inv) {
    const int64_t n = a->n;
    auto out = make_mat(c, n, -1, a->nnz);
    out->diagonal = a->diagonal;
    DBuf<int32_t> len(c, n);
    LB_LAUNCH(c, perm_row_len, cdiv(n, 256), 256, 0, n, a->indptr.p, order, len.p);
    exclusive_scan_i32(c, len.p, out->indptr.p, n);
    LB_LAUNCH(c, perm_copy_rows, cdiv(n * 8, 256), 256, 0, n, a->indptr.p, a->indices.p, a->data.p, order, inv,
              out->indptr.p, out->indices.p, out->data.p);
    return out;
}

// ---- row-partitioned mode: row block with global columns ----------------------------------------
__global__ void shift_ptr_kernel(int64_t nloc, const int32_t *__restrict__ ptr, int32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nloc) out[i] = ptr[i] - ptr[0];
}

std::unique_ptr<lb_mat> row_block(lb_ctx *c, const lb_mat *a, int64_t r0, int64_t r1, int64_t ncols) {
    const int64_t nloc = r1 - r0;
    int32_t ends[2];
    read_back(c, &ends[0], a->indptr.p + r0, 1);
    read_back(c, &ends[1], a->indptr.p + r1, 1);
    const int64_t nnz = ends[1] - ends[0];
    auto out = make_mat(c, nloc, ncols, nnz);
    out->diagonal = false;
    LB_LAUNCH(c, shift_ptr_kernel, cdiv(nloc + 1, 256), 256, 0, nloc, a->indptr.p + r0, out->indptr.p);
    d2d(c, out->indices.p, a->indices.p + ends[0], nnz * sizeof(int32_t));
    d2d(c, out->data.p, a->data.p + ends[0], nnz * sizeof(double));
    return out;
}

__global__ void gather_rows_kernel(int64_t n, int cols, const int32_t *__restrict__ map, const double *__restrict__ x,
                                   int ldx, double *__restrict__ y, int ldy) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t row = t / cols;
    const int col = (int)(t - row * cols);
    y[row * ldy + col] = x[(int64_t)map[row] * ldx + col];
}

// y[i, :] = x[map[i], :]
void gather_rows(lb_ctx *c, int64_t n, int cols, const int32_t *map, const double *x, int ldx, double *y, int ldy) {
    if (n * cols == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 16.0 * n * cols);
    LB_LAUNCH(c, gather_rows_kernel, cdiv(n * cols, 256), 256, 0, n, cols, map, x, ldx, y, ldy);
}

// ---- sorted accumulate into a thread-private slice of global memory ----------------------------
__device__ __forceinline__ void acc_insert(int32_t *keys, double *vals, int &cnt, int key, double v) {
    int pos = cnt;
    while (pos > 0 && keys[pos - 1] >= key) pos--;
    if (pos < cnt && keys[pos] == key) {
        vals[pos] += v;
        return;
    }
    for (int q = cnt; q > pos; q--) {
        keys[q] = keys[q - 1];
        vals[q] = vals[q - 1];
    }
    keys[pos] = key;
    vals[pos] = v;
    cnt++;
}

__global__ void spgemm_ub(int64_t n, const int32_t *__restrict__ aptr, const int32_t *__restrict__ aidx,
                          const int32_t *__restrict__ bptr, int32_t *__restrict__ ub) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int s = 0;
    for (int p = aptr[r]; p < aptr[r + 1]; p++) {
        const int j = aidx[p];
        s += bptr[j + 1] - bptr[j];
    }
    ub[r] = s;
}

__global__ void spgemm_accumulate(int64_t n, const int32_t *__restrict__ aptr, const int32_t *__restrict__ aidx,
                                  const double *__restrict__ aval, const int32_t *__restrict__ bptr,
                                  const int32_t *__restrict__ bidx, const double *__restrict__ bval,
                                  const int32_t *__restrict__ uoff, int32_t *__restrict__ keys,
                                  double *__restrict__ vals, int32_t *__restrict__ cnt_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int32_t *k = keys + uoff[r];
    double *v = vals + uoff[r];
    int cnt = 0;
    for (int p = aptr[r]; p < aptr[r + 1]; p++) {
        const int j = aidx[p];
        const double a = aval[p];
        for (int q = bptr[j]; q < bptr[j + 1]; q++) acc_insert(k, v, cnt, bidx[q], a * bval[q]);
    }
    cnt_out[r] = cnt;
}

__global__ void compact_rows(int64_t n, const int32_t *__restrict__ uoff, const int32_t *__restrict__ keys,
                             const double *__restrict__ vals, const int32_t *__restrict__ cptr,
                             int32_t *__restrict__ cidx, double *__restrict__ cval) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int src = uoff[r], dst = cptr[r], cnt = cptr[r + 1] - dst;
    for (int q = 0; q < cnt; q++) {
        cidx[dst + q] = keys[src + q];
        cval[dst + q] = vals[src + q];
    }
}

static std::unique_ptr<lb_mat> compact_from_scratch(lb_ctx *c, int64_t n, int64_t ncols, const int32_t *uoff,
                                                    const int32_t *keys, const double *vals, const int32_t *cnt) {
    DBuf<int32_t> cptr(c, n + 1);
    exclusive_scan_i32(c, cnt, cptr.p, n);
    int32_t nnz = 0;
    read_back(c, &nnz, cptr.p + n, 1);
    auto out = make_mat(c, n, ncols, nnz);
    d2d(c, out->indptr.p, cptr.p, (n + 1) * sizeof(int32_t));
    LB_LAUNCH(c, compact_rows, cdiv(n, 128), 128, 0, n, uoff, keys, vals, out->indptr.p, out->indices.p, out->data.p);
    return out;
}

std::unique_ptr<lb_mat> spgemm(lb_ctx *c, const lb_mat *a, const lb_mat *b) {
    const int64_t n = a->n;
    const int64_t ncols = b->ncols < 0 ? b->n : b->ncols;
    DBuf<int32_t> ub(c, n), uoff(c, n + 1), cnt(c, n);
    LB_LAUNCH(c, spgemm_ub, cdiv(n, 256), 256, 0, n, a->indptr.p, a->indices.p, b->indptr.p, ub.p);
    exclusive_scan_i32(c, ub.p, uoff.p, n);
    int32_t total = 0;
    read_back(c, &total, uoff.p + n, 1);
    LB_REQUIRE(total >= 0, "SpGEMM intermediate exceeds int32 indexing");
    DBuf<int32_t> keys(c, (size_t)total);
    DBuf<double> vals(c, (size_t)total);
    LB_LAUNCH(c, spgemm_accumulate, cdiv(n, 64), 64, 0, n, a->indptr.p, a->indices.p, a->data.p, b->indptr.p,
              b->indices.p, b->data.p, uoff.p, keys.p, vals.p, cnt.p);
    return compact_from_scratch(c, n, ncols, uoff.p, keys.p, vals.p, cnt.p);
}

// ---- transpose (counting sort by column, then per-row sort for determinism) --------------------
__global__ void count_cols(int64_t nnz, const int32_t *__restrict__ idx, int32_t *__restrict__ cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) atomicAdd(cnt + idx[p], 1);
}

__global__ void transpose_fill(int64_t n, const int32_t *__restrict__ aptr, const int32_t *__restrict__ aidx,
                               const double *__restrict__ aval, const int32_t *__restrict__ tptr,
                               int32_t *__restrict__ cursor, int32_t *__restrict__ tidx, double *__restrict__ tval) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int p = aptr[r]; p < aptr[r + 1]; p++) {
        const int col = aidx[p];
        const int dst = tptr[col] + atomicAdd(cursor + col, 1);
        tidx[dst] = (int)r;
        tval[dst] = aval[p];
    }
}

__global__ void sort_rows(int64_t n, const int32_t *__restrict__ ptr, int32_t *__restrict__ idx,
                          double *__restrict__ val) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int beg = ptr[r], end = ptr[r + 1];
    for (int i = beg + 1; i < end; i++) {
        const int key = idx[i];
        const double v = val[i];
        int j = i - 1;
        while (j >= beg && idx[j] > key) {
            idx[j + 1] = idx[j];
            val[j + 1] = val[j];
            j--;
        }
        idx[j + 1] = key;
        val[j + 1] = v;
    }
}

std::unique_ptr<lb_mat> transpose(lb_ctx *c, const lb_mat *a) {
    const int64_t n = a->n, nc = a->ncols < 0 ? a->n : a->ncols;
    auto t = make_mat(c, nc, n, a->nnz);
    DBuf<int32_t> cnt(c, nc);
    cnt.zero();
    LB_LAUNCH(c, count_cols, cdiv(a->nnz, 256), 256, 0, a->nnz, a->indices.p, cnt.p);
    exclusive_scan_i32(c, cnt.p, t->indptr.p, nc);
    cnt.zero();
    LB_LAUNCH(c, transpose_fill, cdiv(n, 256), 256, 0, n, a->indptr.p, a->indices.p, a->data.p, t->indptr.p, cnt.p,
              t->indices.p, t->data.p);
    LB_LAUNCH(c, sort_rows, cdiv(nc, 128), 128, 0, nc, t->indptr.p, t->indices.p, t->data.p);
    return t;
}

// ---- numbering conversions ---------------------------------------------------------------------------
std::unique_ptr<lb_mat> to_caller_order(lb_ctx *c, const lb_mat *a) {
    LB_REQUIRE(a->permuted && a->ord && a->ord->ready, "matrix is not stored in a locality numbering");
    // caller row i = stored row inv[i]; caller column = order[stored column]
    auto out = permute_symmetric(c, a, a->ord->inv.p, a->ord->order.p);
    LB_LAUNCH(c, sort_rows, cdiv(out->n, 128), 128, 0, out->n, out->indptr.p, out->indices.p, out->data.p);
    return out;
}

static void into_numbering(lb_ctx *c, const lb_mat *x, const std::shared_ptr<lb_order> &ord, MatView &v) {
    if (x->permuted && x->ord == ord) {
        v.m = x;
        return;
    }
    std::unique_ptr<lb_mat> plain;
    const lb_mat *src = x;
    if (x->permuted) {  // assembled on another mesh: through the caller's numbering
        plain = to_caller_order(c, x);
        src = plain.get();
    }
    LB_REQUIRE(src->n == ord->n, "matrix dimensions differ (%lld vs %lld)", (long long)src->n, (long long)ord->n);
    v.owned = permute_symmetric(c, src, ord->order.p, ord->inv.p);
    // sorted rows: a user-assigned matrix with the pattern of an assembled one must match it entry by entry
    LB_LAUNCH(c, sort_rows, cdiv(v.owned->n, 128), 128, 0, v.owned->n, v.owned->indptr.p, v.owned->indices.p,
              v.owned->data.p);
    v.owned->permuted = true;
    v.owned->ord = ord;
    v.m = v.owned.get();
}

std::shared_ptr<lb_order> common_numbering(lb_ctx *c, const lb_mat *a, const lb_mat *b, MatView &va, MatView &vb) {
    std::shared_ptr<lb_order> ord = a->permuted ? a->ord : (b && b->permuted ? b->ord : nullptr);
    if (!ord) {
        va.m = a;
        vb.m = b;
        return nullptr;
    }
    into_numbering(c, a, ord, va);
    if (b) into_numbering(c, b, ord, vb);
    return ord;
}

// =============================================================================================
// aggregation
// =============================================================================================
typedef unsigned long long u64;

__device__ __forceinline__ u64 mis_key(unsigned state, unsigned i) {
    unsigned h = i * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return ((u64)state << 62) | ((u64)(h & 0x3FFFFFFFu) << 32) | (u64)i;
}

__device__ __forceinline__ bool is_strong(double aij, double dii, double djj, double theta2) {
    // |a_ij| >= theta * sqrt(a_ii a_jj)  (theta2 = theta^2; theta = 0: everything is strong)
    return aij * aij >= theta2 * fabs(dii * djj);
}

__global__ void mis_init(int64_t n, int *__restrict__ state, u64 *__restrict__ t) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    state[i] = 1;
    t[i] = mis_key(1, (unsigned)i);
}

__global__ void mis_propagate(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                              const double *__restrict__ val, const double *__restrict__ diag, double theta2,
                              const u64 *__restrict__ tin, u64 *__restrict__ tout) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 best = tin[i];
    const double dii = diag[i];
    for (int p = ptr[i]; p < ptr[i + 1]; p++) {
        const int j = idx[p];
        if (j == i || !is_strong(val[p], dii, diag[j], theta2)) continue;
        const u64 tj = tin[j];
        if (tj > best) best = tj;
    }
    tout[i] = best;
}

__global__ void mis_update(int64_t n, int *__restrict__ state, const u64 *__restrict__ t2, u64 *__restrict__ t0,
                           int *__restrict__ undecided) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = state[i];
    if (s == 1) {
        const u64 t = t2[i];
        if (t == mis_key(1, (unsigned)i)) s = 2;
        else if ((t >> 62) == 2) s = 0;
        else atomicAdd(undecided, 1);
        state[i] = s;
    }
    t0[i] = mis_key((unsigned)s, (unsigned)i);
}

__global__ void flag_roots(int64_t n, const int *__restrict__ state, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = state[i] == 2;
}

__global__ void agg_roots(int64_t n, const int *__restrict__ state, const int32_t *__restrict__ scan,
                          int32_t *__restrict__ agg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) agg[i] = state[i] == 2 ? scan[i] : -1;
}

// unassigned nodes join the aggregate of the strong neighbour with the largest key among those
// already assigned in `agg_in`; roots_only restricts the candidates to MIS roots (pass 1)
__global__ void agg_join(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                         const double *__restrict__ val, const double *__restrict__ diag, double theta2,
                         const int *__restrict__ state, int roots_only, const int32_t *__restrict__ agg_in,
                         int32_t *__restrict__ agg_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = agg_in[i];
    if (a < 0) {
        u64 best = 0;
        const double dii = diag[i];
        for (int p = ptr[i]; p < ptr[i + 1]; p++) {
            const int j = idx[p];
            if (j == i || agg_in[j] < 0 || (roots_only && state[j] != 2)) continue;
            if (!is_strong(val[p], dii, diag[j], theta2)) continue;
            const u64 key = mis_key(1, (unsigned)j);
            if (key > best) {
                best = key;
                a = agg_in[j];
            }
        }
    }
    agg_out[i] = a;
}

__global__ void agg_finish(int64_t n, int32_t *__restrict__ agg, int32_t *__restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = agg[i];
    if (a < 0) agg[i] = a = 0;  // cannot happen for a symmetric strength graph (MIS-2 is maximal)
    atomicAdd(cnt + a, 1);
}

// returns number of aggregates; agg (n) device
static int aggregate(lb_ctx *c, const lb_mat *K, const double *diag, double theta, DBuf<int32_t> &agg,
                     DBuf<int32_t> &agg_cnt) {
    const int64_t n = K->n;
    const double theta2 = theta * theta;
    DBuf<int> state(c, n), undecided(c, 1);
    DBuf<u64> t0(c, n), t1(c, n), t2(c, n);
    const int grid = cdiv(n, 256);
    LB_LAUNCH(c, mis_init, grid, 256, 0, n, state.p, t0.p);
    for (int round = 0; round < 64; round++) {
        LB_LAUNCH(c, mis_propagate, grid, 256, 0, n, K->indptr.p, K->indices.p, K->data.p, diag, theta2, t0.p, t1.p);
        LB_LAUNCH(c, mis_propagate, grid, 256, 0, n, K->indptr.p, K->indices.p, K->data.p, diag, theta2, t1.p, t2.p);
        undecided.zero();
        LB_LAUNCH(c, mis_update, grid, 256, 0, n, state.p, t2.p, t0.p, undecided.p);
        int left = 0;
        read_back(c, &left, undecided.p, 1);
        if (left == 0) break;
    }
    DBuf<int32_t> flag(c, n), scan(c, n + 1), agg1(c, n);
    LB_LAUNCH(c, flag_roots, grid, 256, 0, n, state.p, flag.p);
    exclusive_scan_i32(c, flag.p, scan.p, n);
    int32_t nagg = 0;
    read_back(c, &nagg, scan.p + n, 1);
    agg.alloc(c, n);
    LB_LAUNCH(c, agg_roots, grid, 256, 0, n, state.p, scan.p, agg.p);
    LB_LAUNCH(c, agg_join, grid, 256, 0, n, K->indptr.p, K->indices.p, K->data.p, diag, theta2, state.p, 1, agg.p,
              agg1.p);
    LB_LAUNCH(c, agg_join, grid, 256, 0, n, K->indptr.p, K->indices.p, K->data.p, diag, theta2, state.p, 0, agg1.p,
              agg.p);
    agg_cnt.alloc(c, std::max<int>(nagg, 1));
    agg_cnt.zero();
    LB_LAUNCH(c, agg_finish, grid, 256, 0, n, agg.p, agg_cnt.p);
    return nagg;
}

// =============================================================================================
// prolongator, diagonal, spectral bound, coarse dense
// =============================================================================================
__global__ void diag_inverse(int64_t n, const double *__restrict__ d, double *__restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dinv[i] = d[i] > 0.0 ? 1.0 / d[i] : 0.0;
}

void launch_diag_inverse(lb_ctx *c, int64_t n, const double *d, double *dinv) {
    LB_LAUNCH(c, diag_inverse, cdiv(n, 256), 256, 0, n, d, dinv);
}

__global__ void gershgorin(int64_t n, const int32_t *__restrict__ ptr, const double *__restrict__ val,
                           const double *__restrict__ dinv, u64 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0;
    if (i < n) {
        double s = 0.0;
        for (int p = ptr[i]; p < ptr[i + 1]; p++) s += fabs(val[p]);
        r = s * dinv[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0 && r > 0.0) atomicMax(out, (u64)__double_as_longlong(r));  // r >= 0: bit order == value order
}

// P = (I - omega D^-1 K) T with T[i, agg[i]] = 1/sqrt(|agg|): one thread per row accumulating
// into its slice of a global scratch of size nnz(K) + n
__global__ void smooth_prolongator(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                   const double *__restrict__ val, const double *__restrict__ dinv, double omega,
                                   const int32_t *__restrict__ agg, const int32_t *__restrict__ agg_cnt,
                                   int32_t *__restrict__ keys, double *__restrict__ vals, int32_t *__restrict__ uoff,
                                   int32_t *__restrict__ cnt_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int off = ptr[i] + (int)i;
    uoff[i] = off;
    int32_t *k = keys + off;
    double *v = vals + off;
    int cnt = 0;
    const int ai = agg[i];
    acc_insert(k, v, cnt, ai, rsqrt((double)agg_cnt[ai]));
    const double w = -omega * dinv[i];
    for (int p = ptr[i]; p < ptr[i + 1]; p++) {
        const int aj = agg[idx[p]];
        acc_insert(k, v, cnt, aj, w * val[p] * rsqrt((double)agg_cnt[aj]));
    }
    cnt_out[i] = cnt;
}

__global__ void densify(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                        const double *__restrict__ val, double *__restrict__ dense) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = ptr[i]; p < ptr[i + 1]; p++) dense[i * n + idx[p]] = val[p];
}

__global__ void set_identity(int64_t n, double *__restrict__ dense, int ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dense[i * ld + i] = 1.0;
}

__global__ void fix_empty_diag(int64_t n, double *__restrict__ dense) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dense[i * n + i] <= 0.0) dense[i * n + i] = 1.0;
}

std::unique_ptr<Amg> amg_setup(lb_ctx *c, std::unique_ptr<lb_mat> K0, int mcap, const AmgOptions &opt) {
    auto amg = std::make_unique<Amg>();
    amg->ctx = c;
    amg->cheb_deg = opt.cheb_deg;
    amg->gamma = opt.gamma;
    amg->mcap = mcap;
    cudaEvent_t e0, e1;
    LB_CUDA(cudaEventCreate(&e0));
    LB_CUDA(cudaEventCreate(&e1));
    LB_CUDA(cudaEventRecord(e0, c->stream));
    std::unique_ptr<lb_mat> K = std::move(K0);
    double theta = opt.theta;
    while (true) {
        const int64_t n = K->n;
        amg->levels.emplace_back();
        AmgLevel &L = amg->levels.back();
        DBuf<double> diag(c, n);
        extract_diagonal(c, K.get(), diag.p);
        L.dinv.alloc(c, n);
        LB_LAUNCH(c, diag_inverse, cdiv(n, 256), 256, 0, n, diag.p, L.dinv.p);
        auto finish_dense = [&]() {
            LB_REQUIRE(n <= 20000, "AMG coarsening stalled at %lld unknowns", (long long)n);
            amg->coarse_n = (int)n;
            amg->coarse_chol.alloc(c, (size_t)n * n);
            amg->coarse_chol.zero();
            LB_LAUNCH(c, densify, cdiv(n, 128), 128, 0, n, K->indptr.p, K->indices.p, K->data.p, amg->coarse_chol.p);
            LB_LAUNCH(c, fix_empty_diag, cdiv(n, 128), 128, 0, n, amg->coarse_chol.p);
            dense_chol_solve_prepare(c, (int)n, amg->coarse_chol.p);
            // inverse = solve with the identity (once per hierarchy; symmetric, so row- == column-major)
            amg->coarse_ld = ((int)n + 1) & ~1;
            amg->coarse_inv.alloc(c, (size_t)n * amg->coarse_ld);
            amg->coarse_inv.zero();
            LB_LAUNCH(c, set_identity, cdiv(n, 128), 128, 0, n, amg->coarse_inv.p, amg->coarse_ld);
            dense_chol_solve(c, (int)n, amg->coarse_chol.p, (int)n, amg->coarse_inv.p, amg->coarse_ld);
            L.K = std::move(K);
            L.x.alloc(c, (size_t)n * mcap);
            L.b.alloc(c, (size_t)n * mcap);
        };
        if (n <= opt.max_coarse || (int)amg->levels.size() >= opt.max_levels) {
            finish_dense();
            break;
        }
        DBuf<u64> gmax(c, 1);
        gmax.zero();
        LB_LAUNCH(c, gershgorin, cdiv(n, 256), 256, 0, n, K->indptr.p, K->data.p, L.dinv.p, gmax.p);
        u64 bits = 0;
        read_back(c, &bits, gmax.p, 1);
        double rho;
        std::memcpy(&rho, &bits, 8);
        L.rho = rho > 0.0 ? rho : 2.0;

        DBuf<int32_t> agg, agg_cnt;
        const int nagg = aggregate(c, K.get(), diag.p, theta, agg, agg_cnt);
        theta *= 0.5;
        if (nagg >= n || nagg == 0) {  // no coarsening possible: stop here with a dense solve
            finish_dense();
            break;
        }
        // smoothed prolongator
        {
            DBuf<int32_t> keys(c, (size_t)K->nnz + n), uoff(c, n), cnt(c, n);
            DBuf<double> vals(c, (size_t)K->nnz + n);
            const double omega = (4.0 / 3.0) / L.rho;
            LB_LAUNCH(c, smooth_prolongator, cdiv(n, 128), 128, 0, n, K->indptr.p, K->indices.p, K->data.p, L.dinv.p,
                      omega, agg.p, agg_cnt.p, keys.p, vals.p, uoff.p, cnt.p);
            L.P = compact_from_scratch(c, n, nagg, uoff.p, keys.p, vals.p, cnt.p);
        }
        L.R = transpose(c, L.P.get());
        auto KP = spgemm(c, K.get(), L.P.get());
        auto Kc = spgemm(c, L.R.get(), KP.get());
        Kc->ncols = -1;
        L.K = std::move(K);
        L.r.alloc(c, (size_t)n * mcap);
        L.d.alloc(c, (size_t)n * mcap);
        if (amg->levels.size() > 1) {
            L.x.alloc(c, (size_t)n * mcap);
            L.b.alloc(c, (size_t)n * mcap);
        }
        K = std::move(Kc);
    }
    LB_CUDA(cudaEventRecord(e1, c->stream));
    LB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    amg->setup_ms = ms;
    if (c->trace) {
        fprintf(stderr, "[lb trace] AMG: %zu levels, setup %.2f ms\n", amg->levels.size(), ms);
        for (auto &l : amg->levels)
            fprintf(stderr, "[lb trace]   n=%lld nnz=%lld rho=%.3f\n", (long long)l.K->n, (long long)l.K->nnz, l.rho);
    }
    return amg;
}

// =============================================================================================
// V-cycle / W-cycle, in double or single precision
// =============================================================================================
// The eigensolver applies the cycle in SINGLE precision (amg_apply_f32): a preconditioner only has to
// be a good approximate inverse, LOBPCG's residuals, Gram matrices and Rayleigh-Ritz stay in double.
// Half the bytes per gathered X row and per matrix value -> the cycle's SpMMs run 1.65-1.85x faster
// (profiles/spmm_variants_r2.json).  The linear solves (lb_solve) keep the double-precision cycle.
template <typename T>
__global__ void cheb_first(int64_t n, int m, const T *__restrict__ dinv, const T *__restrict__ src, int ldsrc, T scale,
                           T *__restrict__ d, T *x, int ldx, int zero_guess) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t row = t / m;
    const int col = (int)(t - row * m);
    const T dv = scale * dinv[row] * src[row * ldsrc + col];
    d[row * m + col] = dv;
    T *xp = x + row * ldx + col;
    *xp = zero_guess ? dv : *xp + dv;
}

template <typename T>
__global__ void cheb_d_only(int64_t n, int m, const T *__restrict__ dinv, const T *__restrict__ src, int ldsrc, T scale,
                            T *__restrict__ d) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t row = t / m;
    const int col = (int)(t - row * m);
    d[t] = scale * dinv[row] * src[row * ldsrc + col];
}

template <typename T>
__global__ void cheb_next(int64_t n, int m, const T *__restrict__ dinv, const T *__restrict__ r, T c1, T c2,
                          T *__restrict__ d, T *x, int ldx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t row = t / m;
    const int col = (int)(t - row * m);
    const T dv = c1 * d[t] + c2 * dinv[row] * r[t];
    d[t] = dv;
    x[row * ldx + col] += dv;
}

// x(n, m) = inv(n, n) b(n, m), single precision: the coarsest level of the fp32 cycle (<= 2000 unknowns,
// visited 4 x 3 times per preconditioner application).  CTA = 16 rows x 64 columns, 4 outputs per thread,
// inv / b tiles of 64 k-values staged in shared memory (round 2: the one-output-per-thread version took
// 79 us at 561 x 64 = 3.4 % of a ShapeDNA step).
__global__ void __launch_bounds__(256) coarse_apply_f32(int n, int m, const float *__restrict__ inv, int ldinv,
                                                        const float *__restrict__ b, int ldb, float *__restrict__ x,
                                                        int ldx) {
    __shared__ float s_inv[16][65];
    __shared__ float s_b[64][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int row0 = blockIdx.x * 16, col = blockIdx.y * 64 + tx;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < n; k0 += 64) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int r = i >> 6, k = i & 63;
            s_inv[r][k] = (row0 + r < n && k0 + k < n) ? inv[(size_t)(row0 + r) * ldinv + k0 + k] : 0.f;
        }
        for (int i = threadIdx.x; i < 64 * 64; i += 256) {
            const int k = i >> 6, c2 = i & 63;
            s_b[k][c2] = (k0 + k < n && blockIdx.y * 64 + c2 < m) ? b[(size_t)(k0 + k) * ldb + blockIdx.y * 64 + c2] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 64; k++) {
            const float bv = s_b[k][tx];
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] = fmaf(s_inv[ty * 4 + r][k], bv, acc[r]);
        }
        __syncthreads();
    }
    if (col < m)
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (row0 + ty * 4 + r < n) x[(size_t)(row0 + ty * 4 + r) * ldx + col] = acc[r];
}

__global__ void f32_to_f64_cols(int64_t n, int m, const float *__restrict__ x, int ldx, double *__restrict__ y, int ldy) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const int64_t row = t / m;
    const int col = (int)(t - row * m);
    y[row * ldy + col] = (double)x[row * ldx + col];
}

__global__ void f64_to_f32_cols(int64_t n, int m, int mpad, const double *__restrict__ x, int ldx, float *__restrict__ y,
                                int ldy) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * mpad) return;
    const int64_t row = t / mpad;
    const int col = (int)(t - row * mpad);
    y[row * ldy + col] = col < m ? (float)x[row * ldx + col] : 0.f;
}

void convert_cols_f32(lb_ctx *c, int64_t n, int m, const double *x, int ldx, float *y, int ldy) {
    const int mpad = (m + 3) & ~3;
    if (n * mpad == 0) return;
    ProfScope prof(c, PROF_ELEMENTWISE, 12.0 * n * m);
    LB_LAUNCH(c, f64_to_f32_cols, cdiv(n * mpad, 256), 256, 0, n, m, mpad, x, ldx, y, ldy);
}

// per-type views of a level
template <typename T>
struct LevelView;
template <>
struct LevelView<double> {
    static double *x(AmgLevel &L) { return L.x.p; }
    static double *b(AmgLevel &L) { return L.b.p; }
    static double *r(AmgLevel &L) { return L.r.p; }
    static double *d(AmgLevel &L) { return L.d.p; }
    static const double *dinv(AmgLevel &L) { return L.dinv.p; }
    static void mm(lb_ctx *c, const lb_mat *a, const double *x, int ldx, double *y, int ldy, int m, int mode,
                   const double *b, int ldb, const SpmmEpilogueT<double> *e) {
        spmm(c, a, x, ldx, y, ldy, m, mode, b, ldb, e);
    }
};
template <>
struct LevelView<float> {
    static float *x(AmgLevel &L) { return L.x32.p; }
    static float *b(AmgLevel &L) { return L.b32.p; }
    static float *r(AmgLevel &L) { return L.r32.p; }
    static float *d(AmgLevel &L) { return L.d32.p; }
    static const float *dinv(AmgLevel &L) { return L.dinv32.p; }
    static void mm(lb_ctx *c, const lb_mat *a, const float *x, int ldx, float *y, int ldy, int m, int mode,
                   const float *b, int ldb, const SpmmEpilogueT<float> *e) {
        spmm_f32(c, a, x, ldx, y, ldy, m, mode, b, ldb, e);
    }
};

// returns true when the result was written to out64 (fused exit of the single-precision cycle)
template <typename T>
static bool smooth(Amg &amg, int l, T *x, int ldx, const T *b, int ldb, int m, bool zero_guess, double *out64 = nullptr,
                   int ldout64 = 0) {
    typedef LevelView<T> LV;
    lb_ctx *c = amg.ctx;
    AmgLevel &L = amg.levels[l];
    const int64_t n = L.K->n;
    const double hi = L.rho, lo = L.rho / 8.0;
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    double rho_k = 1.0 / sigma;
    const int grid = cdiv(n * m, 256);
    T *Lr = LV::r(L), *Ld = LV::d(L);
    const T *dinv = LV::dinv(L);
    if (amg.cheb_deg == 2 && m > 2 && !L.K->diagonal) {
        // fused form: the elementwise Chebyshev updates ride in the SpMM epilogues (36 % less HBM
        // traffic per cycle on the finest level than the separate kernels below)
        const double rho_n = 1.0 / (2.0 * sigma - rho_k);
        SpmmEpilogueT<T> e{};
        e.dinv = dinv;
        const T *src = b;
        int ldsrc = ldb;
        if (zero_guess) {
            ProfScope prof(c, PROF_ELEMENTWISE, 2.0 * sizeof(T) * n * m);
            LB_LAUNCH(c, cheb_d_only<T>, grid, 256, 0, n, m, dinv, b, ldb, (T)(1.0 / theta), Ld);  // d = dinv o b / theta
        } else {
            e.out2 = Ld;
            e.ldout2 = m;
            e.c2 = (T)(1.0 / theta);
            LV::mm(c, L.K.get(), x, ldx, Lr, m, m, 3, b, ldb, &e);  // r = b - K x, d = dinv o r / theta
            src = Lr;
            ldsrc = m;
        }
        e.out2 = x;
        e.ldout2 = ldx;
        e.c1 = (T)(rho_n * rho_k);
        e.c2 = (T)(2.0 * rho_n / delta);
        e.overwrite = zero_guess ? 1 : 0;
        e.out64 = out64;
        e.ldout64 = ldout64;
        LV::mm(c, L.K.get(), Ld, m, nullptr, 0, m, 4, src, ldsrc, &e);  // x (+)= d + [c1 d + c2 dinv o (src - K d)]
        return out64 != nullptr;
    }
    const T *src = b;
    int ldsrc = ldb;
    if (!zero_guess) {
        LV::mm(c, L.K.get(), x, ldx, Lr, m, m, 1, b, ldb, nullptr);  // r = b - K x
        src = Lr;
        ldsrc = m;
    }
    {
        ProfScope prof(c, PROF_ELEMENTWISE, (zero_guess ? 3.0 : 4.0) * sizeof(T) * n * m);
        LB_LAUNCH(c, cheb_first<T>, grid, 256, 0, n, m, dinv, src, ldsrc, (T)(1.0 / theta), Ld, x, ldx, (int)zero_guess);
    }
    for (int k = 1; k < amg.cheb_deg; k++) {
        LV::mm(c, L.K.get(), Ld, m, Lr, m, m, 1, src, ldsrc, nullptr);  // r = r_prev - K d
        src = Lr;
        ldsrc = m;
        const double rho_n = 1.0 / (2.0 * sigma - rho_k);
        {
            ProfScope prof(c, PROF_ELEMENTWISE, 5.0 * sizeof(T) * n * m);
            LB_LAUNCH(c, cheb_next<T>, grid, 256, 0, n, m, dinv, Lr, (T)(rho_n * rho_k), (T)(2.0 * rho_n / delta), Ld, x, ldx);
        }
        rho_k = rho_n;
    }
    return false;
}

static void coarse_solve(Amg &amg, AmgLevel &L, const double *b, int ldb, double *x, int ldx, int m) {
    lb_ctx *c = amg.ctx;
    const bool aligned = (ldb % 2 == 0) && ((reinterpret_cast<uintptr_t>(b) & 15) == 0) && b != x;
    if (aligned) {  // x = K^-1 b as one tall-skinny DMMA product with the explicit inverse
        ProfScope prof(c, PROF_TRSM, 2.0 * amg.coarse_n * amg.coarse_n * m, amg.coarse_n, m);
        update_dmma(c, amg.coarse_n, amg.coarse_n, amg.coarse_inv.p, amg.coarse_ld, m, b, ldb, 1.0, 0.0, x, ldx);
        return;
    }
    copy_cols(c, L.K->n, m, b, ldb, x, ldx);
    dense_chol_solve(c, amg.coarse_n, amg.coarse_chol.p, m, x, ldx);
}

static void coarse_solve(Amg &amg, AmgLevel &, const float *b, int ldb, float *x, int ldx, int m) {
    lb_ctx *c = amg.ctx;
    ProfScope prof(c, PROF_TRSM, 2.0 * amg.coarse_n * amg.coarse_n * m, amg.coarse_n, kProfF32 + m);
    LB_LAUNCH(c, coarse_apply_f32, dim3(cdiv(amg.coarse_n, 16), cdiv(m, 64)), 256, 0, amg.coarse_n, m, amg.coarse_inv32.p,
              amg.coarse_ld32, b, ldb, x, ldx);
}

// one multigrid cycle on level l: x <- x + cycle(b - K x) (x = 0 on entry when zero_guess).
// amg.gamma = 1: V-cycle; 2: W-cycle (the coarse problem is visited twice - cheap, levels shrink
// ~10x - which keeps the convergence factor level-independent for MIS-2 aggregates)
template <typename T>
static bool cycle(Amg &amg, int l, const T *b, int ldb, T *x, int ldx, int m, bool zero_guess, double *out64 = nullptr,
                  int ldout64 = 0) {
    typedef LevelView<T> LV;
    lb_ctx *c = amg.ctx;
    AmgLevel &L = amg.levels[l];
    if (l == (int)amg.levels.size() - 1) {
        coarse_solve(amg, L, b, ldb, x, ldx, m);
        return false;
    }
    AmgLevel &C = amg.levels[l + 1];
    smooth<T>(amg, l, x, ldx, b, ldb, m, zero_guess);
    LV::mm(c, L.K.get(), x, ldx, LV::r(L), m, m, 1, b, ldb, nullptr);          // r = b - K x
    LV::mm(c, L.R.get(), LV::r(L), m, LV::b(C), m, m, 0, nullptr, 0, nullptr);  // b_c = R r
    const bool coarsest_next = l + 1 == (int)amg.levels.size() - 1;
    const int visits = coarsest_next ? 1 : amg.gamma;
    for (int g = 0; g < visits; g++) cycle<T>(amg, l + 1, LV::b(C), m, LV::x(C), m, m, g == 0);
    LV::mm(c, L.P.get(), LV::x(C), m, x, ldx, m, 2, x, ldx, nullptr);  // x += P x_c
    return smooth<T>(amg, l, x, ldx, b, ldb, m, false, out64, ldout64);
}

void amg_apply(Amg &amg, const double *r, int ldr, double *z, int ldz, int m, int level) {
    LB_REQUIRE(m <= amg.mcap, "AMG applied to %d columns but sized for %d", m, amg.mcap);
    cycle<double>(amg, level, r, ldr, z, ldz, m, true);
}

__global__ void to_f32_vec(int64_t n, const double *__restrict__ in, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// single-precision mirrors of the hierarchy (values, diagonals, coarse inverse, work blocks): made on
// the first single-precision application.  Returns false when a level does not fit the strip SpMM.
bool amg_prepare_f32(Amg &amg) {
    if (amg.f32_state) return amg.f32_state > 0;
    lb_ctx *c = amg.ctx;
    amg.f32_state = -1;
    if (amg.cheb_deg != 2 || amg.levels.size() < 2) return false;
    const int nl = (int)amg.levels.size();
    for (int l = 0; l < nl; l++) {
        AmgLevel &L = amg.levels[l];
        if (l + 1 < nl && (!spmm_f32_supported(c, L.K.get()) || !spmm_f32_supported(c, L.P.get()) ||
                           !spmm_f32_supported(c, L.R.get())))
            return false;
    }
    const int mc = (amg.mcap + 3) & ~3;
    for (int l = 0; l < nl; l++) {
        AmgLevel &L = amg.levels[l];
        const int64_t n = L.K->n;
        L.dinv32.alloc(c, n);
        LB_LAUNCH(c, to_f32_vec, cdiv(n, 256), 256, 0, n, L.dinv.p, L.dinv32.p);
        L.x32.alloc(c, (size_t)n * mc);
        L.b32.alloc(c, (size_t)n * mc);
        if (l + 1 < nl) {
            L.r32.alloc(c, (size_t)n * mc);
            L.d32.alloc(c, (size_t)n * mc);
            mat_values_f32(c, L.K.get());
            mat_values_f32(c, L.P.get());
            mat_values_f32(c, L.R.get());
        }
    }
    amg.coarse_ld32 = (amg.coarse_n + 3) & ~3;
    amg.coarse_inv32.alloc(c, (size_t)amg.coarse_n * amg.coarse_ld32);
    amg.coarse_inv32.zero();
    f64_to_f32_cols<<<cdiv((int64_t)amg.coarse_n * amg.coarse_ld32, 256), 256, 0, c->stream>>>(
        amg.coarse_n, amg.coarse_n, amg.coarse_ld32, amg.coarse_inv.p, amg.coarse_ld, amg.coarse_inv32.p, amg.coarse_ld32);
    c->launches++;
    LB_CUDA(cudaGetLastError());
    amg.f32_state = 1;
    return true;
}

void amg_apply_f32(Amg &amg, const float *r, int ldr, double *z, int ldz, int m, int mz, int level) {
    LB_REQUIRE(amg.f32_state > 0, "single-precision hierarchy not prepared");
    LB_REQUIRE(m % 4 == 0 && m <= ((amg.mcap + 3) & ~3) && ldr % 4 == 0 && mz <= m, "single-precision cycle: bad block shape");
    AmgLevel &L = amg.levels[level];
    // the fused exit writes all m (padded) columns: only when the caller's rows have room for them
    const bool direct = (mz == m || ldz >= m) && (ldz % 2 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
    // kF32Cycles cycles per application: z_{k+1} = z_k + cycle(r - K z_k).  At ~3.5 ms per cycle (level 9,
    // 64 columns) against ~22 ms for the rest of a LOBPCG iteration a stronger preconditioner pays:
    // measured 40 / 31 / 27 iterations and 1117 / 938 / 885 ms for 1 / 2 / 3 cycles (level-9 icosphere),
    // 47 / 34 / 28 iterations and 1152 / 977 / 927 ms on the 121^3 tet cube
    // (V-cycles instead of W-cycles: 37 / 33 / 31 iterations with 3 / 4 / 5 of them, all slower in total)
    constexpr int ncyc = kF32Cycles;
    for (int k = 0; k + 1 < ncyc; k++) cycle<float>(amg, level, r, ldr, L.x32.p, m, m, k == 0);
    const bool done = cycle<float>(amg, level, r, ldr, L.x32.p, m, m, ncyc == 1, direct ? z : nullptr, ldz);
    if (!done) {
        ProfScope prof(amg.ctx, PROF_ELEMENTWISE, 12.0 * L.K->n * mz);
        LB_LAUNCH(amg.ctx, f32_to_f64_cols, cdiv(L.K->n * mz, 256), 256, 0, L.K->n, mz, L.x32.p, m, z, ldz);
    }
}

}  // namespace lb
