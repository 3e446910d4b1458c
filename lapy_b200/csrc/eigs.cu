// eigs.cu - block LOBPCG on the pencil (A, B) with an AMG V-cycle preconditioner (lb_eigs).
//
// Replaces Solver.eigs (lapy/solver.py:667-716): splu(A - sigma*B) + ARPACK shift-invert Lanczos.
// For sigma <= 0 "the k eigenvalues nearest sigma" are the k smallest, which is what a
// preconditioned block eigensolver delivers; the reference's shift enters as the SPD operator
// K = A - sigma*B the preconditioner approximates the inverse of (the same operator SuperLU
// factorises at solver.py:707).
//
// Algorithm (Knyazev's LOBPCG in the robust basis form of Duersch/Shao/Yang/Gu 2018):
//   S = [X | P | W] kept B-orthonormal; W = B-orthonormalised, X- and P-orthogonalised
//   preconditioned residuals of the not yet converged columns (soft locking); Rayleigh-Ritz on
//   S^T A S (<= 3m x 3m, cuSOLVER syevd); new X = S Cx, new P = S Cp with Cp the W/P part of Cx
//   re-orthonormalised against Cx in coefficient space.  All n-sized work is SpMM / tall-skinny
//   block products on the device; the host only handles <= 3m x 3m coefficient matrices.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include <cusolverDn.h>

#include <map>

#include "amg.cuh"
#include "dist.cuh"

namespace lb {

// ---- row-partitioned mode: the three operations that communicate --------------------------------
// Every rank owns the rows [rank*rpr, min(n, (rank+1)*rpr)) of the renumbered operator and the
// matching rows of all block vectors.  SpMM exchanges the boundary rows of the block vector with
// the neighbouring ranks (grouped ncclSend / ncclRecv); Gram matrices and column dots are summed
// over the ranks (ncclAllReduce).  With D == NULL these are the single-GPU operations.
using RawBuf = RawBufT<double>;  // plain cudaMalloc (dist.cuh): NCCL must not be handed pool memory

// Halo exchange of the row-partitioned SpMM.  Per operator: the ghost columns of this rank's row block (sorted
// global ids, grouped by owner), the rows every peer needs from this rank, and the row block
// with its columns renumbered to [own rows | ghost rows], so that the ordinary SpMM kernel runs
// on one contiguous (n_loc + n_ghost, w) input whose ghost part is filled by grouped
// ncclSend / ncclRecv.  For a Morton-ordered surface the ghosts are O(sqrt(n)) rows per
// neighbour instead of the n rows the all-gather moves.
struct HaloPlan {
    int64_t n_loc = 0, n_ghost = 0, n_send = 0;
    int wcap = 0;
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt;  // rows, per peer
    RawBufT<int32_t> send_rows;     // local index of every row to send, grouped by destination
    std::unique_ptr<lb_mat> local;  // row block, columns renumbered [own | ghosts] (unsorted rows)
    RawBuf sendbuf, xg;             // (n_send, wcap) packed rows; (n_loc + n_ghost, wcap) SpMM input
};

struct DistOps {
    const DistCtx *d = nullptr;
    int64_t rpr = 0;        // rows per rank (last rank may own fewer)
    int64_t n_local = 0;
    int64_t r0 = 0, r1 = 0;  // this rank's rows
    RawBuf pack, gath, red;  // (rpr, k), (world*rpr, k): final eigenvector all-gather; all-reduce staging
    int wcap = 0;
    std::map<const lb_mat *, std::unique_ptr<HaloPlan>> plans;  // built on first use (collectively)
    // replicated preconditioner: every rank holds the AMG hierarchy of the FULL operator and applies
    // it to its share of the COLUMNS; see dist_precond
    Amg *full_amg = nullptr;
    int64_t n_full = 0;
    RawBuf tpack, tfull_r, tfull_z;  // (n_local, mcap) packed slices; (n, ceil(mcap/world)) in / out
};

__global__ void halo_remap_kernel(int64_t nnz, const int32_t *__restrict__ idx, int r0, int r1, int nloc,
                                  const int32_t *__restrict__ ghost, int ng, int32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int col = idx[i];
    if (col >= r0 && col < r1) {
        out[i] = col - r0;
        return;
    }
    int lo = 0, hi = ng;  // first position with ghost[pos] >= col (col is in the list)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ghost[mid] < col) lo = mid + 1;
        else hi = mid;
    }
    out[i] = nloc + lo;
}

__global__ void halo_shift_kernel(int64_t cnt, int32_t *rows, int r0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) rows[i] -= r0;
}

// collective over the communicator: every rank calls it for its own row block of the same operator
static std::unique_ptr<HaloPlan> build_halo(lb_ctx *c, const DistCtx *d, const lb_mat *rows, int64_t r0, int64_t r1,
                                            int64_t rpr, int wcap) {
    const int W = d->world, me = d->rank;
    auto h = std::make_unique<HaloPlan>();
    h->n_loc = r1 - r0;
    h->wcap = wcap;
    const int64_t nnz = rows->nnz;
    std::vector<int32_t> hidx((size_t)nnz);
    d2h(c, hidx.data(), rows->indices.p, (size_t)nnz * sizeof(int32_t));
    sync(c);
    std::vector<uint8_t> mark((size_t)W * rpr, 0);
    for (int32_t col : hidx)
        if (col < r0 || col >= r1) mark[(size_t)col] = 1;
    std::vector<int32_t> ghosts;
    for (size_t i = 0; i < mark.size(); i++)
        if (mark[i]) ghosts.push_back((int32_t)i);
    h->n_ghost = (int64_t)ghosts.size();
    h->recv_cnt.assign(W, 0);
    h->recv_off.assign(W, 0);
    h->send_cnt.assign(W, 0);
    h->send_off.assign(W, 0);
    for (int32_t g : ghosts) h->recv_cnt[g / rpr]++;
    for (int p = 1; p < W; p++) h->recv_off[p] = h->recv_off[p - 1] + h->recv_cnt[p - 1];
    // who needs how many of my rows: all-gather the request counts
    RawBufT<int32_t> cnt_loc, cnt_all, d_ghost;
    cnt_loc.alloc(W);
    cnt_all.alloc((size_t)W * W);
    std::vector<int32_t> hc(W), hall((size_t)W * W);
    for (int p = 0; p < W; p++) hc[p] = (int32_t)h->recv_cnt[p];
    h2d(c, cnt_loc.p, hc.data(), W * sizeof(int32_t));
    dist_allgather_i32(c, d, cnt_loc.p, cnt_all.p, W);
    d2h(c, hall.data(), cnt_all.p, hall.size() * sizeof(int32_t));
    sync(c);
    for (int p = 0; p < W; p++) h->send_cnt[p] = p == me ? 0 : hall[(size_t)p * W + me];
    for (int p = 1; p < W; p++) h->send_off[p] = h->send_off[p - 1] + h->send_cnt[p - 1];
    h->n_send = h->send_off[W - 1] + h->send_cnt[W - 1];
    // request lists: my ghost ids go to their owners, the owners' requests come back as global row ids
    d_ghost.alloc(std::max<size_t>(1, ghosts.size()));
    if (!ghosts.empty()) h2d(c, d_ghost.p, ghosts.data(), ghosts.size() * sizeof(int32_t));
    h->send_rows.alloc(std::max<size_t>(1, (size_t)h->n_send));
    dist_exchange(c, d, d_ghost.p, h->recv_off.data(), h->recv_cnt.data(), h->send_rows.p, h->send_off.data(),
                  h->send_cnt.data(), 4);
    if (h->n_send) {
        LB_LAUNCH(c, halo_shift_kernel, cdiv(h->n_send, 256), 256, 0, h->n_send, h->send_rows.p, (int)r0);
        std::vector<int32_t> chk((size_t)h->n_send);
        d2h(c, chk.data(), h->send_rows.p, chk.size() * sizeof(int32_t));
        sync(c);
        for (int32_t v : chk) LB_REQUIRE(v >= 0 && v < h->n_loc, "halo plan: peer requested row %d outside this rank's block", v);
    }
    // the row block with columns renumbered [own | ghosts]
    auto m = std::make_unique<lb_mat>();
    m->ctx = c;
    m->n = h->n_loc;
    m->ncols = h->n_loc + h->n_ghost;
    m->nnz = nnz;
    m->indptr.alloc(c, h->n_loc + 1);
    m->indices.alloc(c, nnz);
    m->data.alloc(c, nnz);
    d2d(c, m->indptr.p, rows->indptr.p, (size_t)(h->n_loc + 1) * sizeof(int32_t));
    d2d(c, m->data.p, rows->data.p, (size_t)nnz * sizeof(double));
    if (nnz)
        LB_LAUNCH(c, halo_remap_kernel, cdiv(nnz, 256), 256, 0, nnz, rows->indices.p, (int)r0, (int)r1, (int)h->n_loc,
                  d_ghost.p, (int)h->n_ghost, m->indices.p);
    h->local = std::move(m);
    h->sendbuf.alloc(std::max<size_t>(1, (size_t)h->n_send * wcap));
    h->xg.alloc((size_t)(h->n_loc + h->n_ghost) * wcap);
    sync(c);  // d_ghost and the host vectors go out of scope
    if (c->trace)
        fprintf(stderr, "[lb trace] rank %d: halo plan: %lld own rows, %lld ghost rows, %lld rows to send\n", me,
                (long long)h->n_loc, (long long)h->n_ghost, (long long)h->n_send);
    return h;
}

static void d_allreduce(lb_ctx *c, DistOps *D, double *buf, size_t count) {
    LB_REQUIRE(count <= D->red.n, "all-reduce staging buffer too small");
    d2d(c, D->red.p, buf, count * sizeof(double));
    dist_allreduce_sum(c, D->d, D->red.p, count);
    d2d(c, buf, D->red.p, count * sizeof(double));
}

static void d_spmm(lb_ctx *c, DistOps *D, const lb_mat *a, const double *x, int ldx, double *y, int ldy, int w) {
    if (!D) {
        spmm(c, a, x, ldx, y, ldy, w);
        return;
    }
    auto &slot = D->plans[a];
    if (!slot) slot = build_halo(c, D->d, a, D->r0, D->r1, D->rpr, D->wcap);
    HaloPlan &h = *slot;
    LB_REQUIRE(w <= h.wcap, "halo exchange: block of %d columns exceeds the plan's capacity %d", w, h.wcap);
    const int W = D->d->world;
    copy_cols(c, h.n_loc, w, x, ldx, h.xg.p, w);
    if (h.n_send) gather_rows(c, h.n_send, w, h.send_rows.p, x, ldx, h.sendbuf.p, w);
    std::vector<int64_t> so(W), sc(W), ro(W), rc(W);
    for (int p = 0; p < W; p++) {
        so[p] = h.send_off[p] * w;
        sc[p] = h.send_cnt[p] * w;
        ro[p] = (h.n_loc + h.recv_off[p]) * w;
        rc[p] = h.recv_cnt[p] * w;
    }
    dist_exchange(c, D->d, h.sendbuf.p, so.data(), sc.data(), h.xg.p, ro.data(), rc.data(), 8);
    spmm(c, h.local.get(), h.xg.p, w, y, ldy, w);
}

static void d_gram(lb_ctx *c, DistOps *D, int64_t n, int p, const double *x, int ldx, int q, const double *y, int ldy,
                   double *cmat, bool symmetric = false) {
    gram(c, n, p, x, ldx, q, y, ldy, cmat, symmetric);
    if (D) d_allreduce(c, D, cmat, (size_t)p * q);
}

static void d_dots(lb_ctx *c, DistOps *D, int64_t n, int cols, const double *x, int ldx, const double *y, int ldy,
                   double *out) {
    col_dots(c, n, cols, x, ldx, y, ldy, out);
    if (D) d_allreduce(c, D, out, cols);
}

// Preconditioner of the row-partitioned mode with a replicated hierarchy: the W-cycle of the full
// operator is applied column-wise in parallel.  Rank p takes the columns [p*mc, (p+1)*mc) of the
// residual block: an all-to-all (grouped send/recv) turns the row-partitioned (n_local, ma) block
// into a column-partitioned (n, mc) one, every rank runs the same cycle the single-GPU solver runs
// (so the iteration count does not grow with the number of ranks, unlike the block-Jacobi
// hierarchy), and the reverse all-to-all brings the rows back.  Traffic per application and rank:
// 2 * n_local * ma * (world-1)/world doubles over NVLink, no communication inside the cycle.
static void dist_precond(lb_ctx *c, DistOps *D, const double *r, int ldr, double *z, int ldz, int ma) {
    const int W = D->d->world, me = D->d->rank;
    const int mc = (ma + W - 1) / W;
    auto c0 = [&](int p) { return std::min(ma, p * mc); };
    auto cw = [&](int p) { return std::min(ma, (p + 1) * mc) - c0(p); };
    auto row0 = [&](int p) { return std::min(D->n_full, (int64_t)p * D->rpr); };
    auto rows = [&](int p) { return std::min(D->n_full, (int64_t)(p + 1) * D->rpr) - row0(p); };
    const int64_t nl = D->n_local;
    const int mine = cw(me);
    LB_REQUIRE((size_t)nl * ma <= D->tpack.n && (size_t)D->n_full * mc <= D->tfull_r.n, "dist_precond: staging buffers too small");
    std::vector<int64_t> so(W), sc(W), ro(W), rc(W);
    // forward: my rows of column slice p -> rank p; rank p's rows of MY slice -> rows [row0(p), ...) of tfull_r
    for (int p = 0; p < W; p++) {
        so[p] = nl * c0(p);
        sc[p] = nl * cw(p);
        ro[p] = row0(p) * mine;
        rc[p] = rows(p) * mine;
        if (cw(p) == 0) continue;
        if (p == me) copy_cols(c, nl, mine, r + c0(p), ldr, D->tfull_r.p + ro[p], mine);
        else copy_cols(c, nl, cw(p), r + c0(p), ldr, D->tpack.p + so[p], cw(p));
    }
    dist_exchange(c, D->d, D->tpack.p, so.data(), sc.data(), D->tfull_r.p, ro.data(), rc.data(), 8);
    if (mine && amg_prepare_f32(*D->full_amg)) {  // single-precision cycle, like the single-GPU solver
        const int mpad = (mine + 3) & ~3;
        DBuf<float> rf(c, (size_t)D->n_full * mpad);
        convert_cols_f32(c, D->n_full, mine, D->tfull_r.p, mine, rf.p, mpad);
        amg_apply_f32(*D->full_amg, rf.p, mpad, D->tfull_z.p, mine, mpad, mine, 0);
    } else if (mine) {
        amg_apply(*D->full_amg, D->tfull_r.p, mine, D->tfull_z.p, mine, mine, 0);
    }
    // reverse: rows [row0(p), ...) of my result slice -> rank p; rank p's slice of MY rows -> tpack -> z
    dist_exchange(c, D->d, D->tfull_z.p, ro.data(), rc.data(), D->tpack.p, so.data(), sc.data(), 8);
    for (int p = 0; p < W; p++) {
        if (cw(p) == 0) continue;
        if (p == me) copy_cols(c, nl, mine, D->tfull_z.p + ro[p], mine, z + c0(p), ldz);
        else copy_cols(c, nl, cw(p), D->tpack.p + so[p], cw(p), z + c0(p), ldz);
    }
}

__global__ void set_column(int64_t n, double *x, int ld, int col, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i * ld + col] = v;
}

// Cholesky-QR of W (n, q) in the B inner product, repeated twice; falls back to an eigen-based
// whitening (SVQB) when the Gram matrix is numerically singular.  On return bw = B w.
// Returns the number of columns kept (columns with negligible norm are dropped by SVQB).
static int b_orthonormalize(lb_ctx *c, DistOps *D, const lb_mat *B, int64_t n, int q, double *w, int ldw, double *bw,
                            int ldbw, double *tmp /* n x q scratch, ld = q */, bool need_bw = true) {
    DBuf<double> G(c, (size_t)q * q), ev(c, q);
    int kept = q;
    for (int rep = 0; rep < 2; rep++) {
        d_spmm(c, D, B, w, ldw, bw, ldbw, kept);
        d_gram(c, D, n, kept, w, ldw, kept, bw, ldbw, G.p, true);
        // Cholesky of the diagonally scaled Gram matrix G' = D G D (D = diag(G)^-1/2) on the HOST: q <= m,
        // O(q^3 / 3) flops, identical on every rank (G is all-reduced).  The columns of W = T(residual)
        // differ in norm by orders of magnitude; without the scaling the pivot ratio below measured the
        // column scaling instead of the conditioning, so nearly every iteration paid a second pass.
        std::vector<double> hL((size_t)kept * kept), hT((size_t)kept * kept, 0.0), dsc(kept);
        d2h(c, hL.data(), G.p, hL.size() * sizeof(double));
        sync(c);
        int info = 0;
        for (int i = 0; i < kept; i++) {
            const double gii = hL[(size_t)i * kept + i];
            if (!(gii > 0.0) || !std::isfinite(gii)) info = i + 1;
            dsc[i] = info ? 1.0 : 1.0 / std::sqrt(gii);
        }
        double dmin = 1e300, dmax = 0;
        std::vector<double> hGs;  // the scaled matrix, kept for the whitening fallback
        if (info == 0) {
            for (int i = 0; i < kept; i++)
                for (int j = 0; j < kept; j++) hL[(size_t)i * kept + j] *= dsc[i] * dsc[j];
            hGs = hL;
            for (int j = 0; j < kept && info == 0; j++) {  // row-major lower Cholesky, in place
                double djj = hL[(size_t)j * kept + j];
                for (int t = 0; t < j; t++) djj -= hL[(size_t)j * kept + t] * hL[(size_t)j * kept + t];
                if (!(djj > 1e-12)) {  // cond(G') > ~1e12: whitening instead
                    info = j + 1;
                    break;
                }
                djj = std::sqrt(djj);
                hL[(size_t)j * kept + j] = djj;
                dmin = std::min(dmin, djj);
                dmax = std::max(dmax, djj);
                for (int i = j + 1; i < kept; i++) {
                    double v = hL[(size_t)i * kept + j];
                    for (int t = 0; t < j; t++) v -= hL[(size_t)i * kept + t] * hL[(size_t)j * kept + t];
                    hL[(size_t)i * kept + j] = v / djj;
                }
            }
        }
        if (info == 0) {
            // W <- W D L'^-T as ONE block update with the explicit (kept x kept) matrix T = D L'^-T
            std::vector<double> Li((size_t)kept * kept, 0.0);
            for (int j = 0; j < kept; j++) {
                Li[(size_t)j * kept + j] = 1.0 / hL[(size_t)j * kept + j];
                for (int i = j + 1; i < kept; i++) {
                    double sacc = 0;
                    for (int t = j; t < i; t++) sacc += hL[(size_t)i * kept + t] * Li[(size_t)t * kept + j];
                    Li[(size_t)i * kept + j] = -sacc / hL[(size_t)i * kept + i];
                }
            }
            for (int i = 0; i < kept; i++)
                for (int j = 0; j <= i; j++) hT[(size_t)j * kept + i] = dsc[j] * Li[(size_t)i * kept + j];
            DBuf<double> dT(c, hT.size());
            h2d(c, dT.p, hT.data(), hT.size() * sizeof(double));
            bool inplace = false;
            {
                ProfScope prof(c, PROF_UPDATE, 2.0 * n * kept * kept, kept, kept);
                inplace = update_dmma_inplace(c, n, kept, w, ldw, dT.p, kept);
            }
            if (!inplace) {
                update(c, n, kept, w, ldw, kept, dT.p, kept, 1.0, 0.0, tmp, q);
                copy_cols(c, n, kept, tmp, q, w, ldw);
            }
            sync(c);
            // cond(G) ~ (dmax/dmin)^2: one pass leaves an orthogonality error ~ eps*cond(G)
            if (rep == 0 && (dmax / dmin) * (dmax / dmin) < 1e3) break;
            continue;
        }
        // SVQB on the scaled matrix: G' = V diag(e) V^T; W <- W D V diag(e)^-1/2 over the columns with e > eps * e_max
        std::vector<double> hV((size_t)kept * kept), hE(kept);
        if (!hGs.empty()) h2d(c, G.p, hGs.data(), hGs.size() * sizeof(double));  // whiten the scaled matrix (else dsc = 1)
        info = sym_eig(c, kept, G.p, ev.p);
        LB_REQUIRE(info == 0, "whitening eigen-decomposition failed (info=%d)", info);
        d2h(c, hV.data(), G.p, hV.size() * sizeof(double));
        d2h(c, hE.data(), ev.p, kept * sizeof(double));
        sync(c);
        const double emax = std::max(hE[kept - 1], 0.0);
        std::vector<double> T;  // (kept, newq) row-major
        int newq = 0;
        for (int j = kept - 1; j >= 0; j--)
            if (hE[j] > 1e-13 * emax && hE[j] > 0.0) newq++;
        if (newq == 0) return 0;
        T.assign((size_t)kept * newq, 0.0);
        int col = 0;
        for (int j = kept - 1; j >= 0 && col < newq; j--, col++) {
            const double s = 1.0 / std::sqrt(hE[j]);
            for (int i = 0; i < kept; i++) T[(size_t)i * newq + col] = dsc[i] * hV[(size_t)j * kept + i] * s;  // row j = evec j
        }
        DBuf<double> dT(c, T.size());
        h2d(c, dT.p, T.data(), T.size() * sizeof(double));
        update(c, n, kept, w, ldw, newq, dT.p, newq, 1.0, 0.0, tmp, q);
        copy_cols(c, n, newq, tmp, q, w, ldw);
        sync(c);
        kept = newq;
    }
    if (need_bw) d_spmm(c, D, B, w, ldw, bw, ldbw, kept);
    return kept;
}

// ---- host-side small dense helpers (column-major-free: everything row-major) ------------------
// orthonormalise the columns of Q (s x q, row-major) against the orthonormal columns of C
// (s x m) and among themselves (modified Gram-Schmidt, twice); drops dependent columns
static int host_orth(int s, int m, const std::vector<double> &C, int q, std::vector<double> &Q) {
    std::vector<double> col(s);
    int kept = 0;
    for (int j = 0; j < q; j++) {
        for (int i = 0; i < s; i++) col[i] = Q[(size_t)i * q + j];
        double n0 = 0;
        for (int i = 0; i < s; i++) n0 += col[i] * col[i];
        n0 = std::sqrt(n0);
        if (!(n0 > 0)) continue;
        for (int rep = 0; rep < 2; rep++) {
            for (int k = 0; k < m; k++) {
                double d = 0;
                for (int i = 0; i < s; i++) d += C[(size_t)i * m + k] * col[i];
                for (int i = 0; i < s; i++) col[i] -= d * C[(size_t)i * m + k];
            }
            for (int k = 0; k < kept; k++) {
                double d = 0;
                for (int i = 0; i < s; i++) d += Q[(size_t)i * q + k] * col[i];
                for (int i = 0; i < s; i++) col[i] -= d * Q[(size_t)i * q + k];
            }
        }
        double n1 = 0;
        for (int i = 0; i < s; i++) n1 += col[i] * col[i];
        n1 = std::sqrt(n1);
        if (n1 < 1e-8 * n0 || n1 < 1e-14) continue;
        for (int i = 0; i < s; i++) Q[(size_t)i * q + kept] = col[i] / n1;
        kept++;
    }
    return kept;
}

// Rayleigh-Ritz coefficient matrix on the device (one CTA, everything in shared memory):
//   Cx  = the m lowest eigenvectors of S^T A S (rows of evT),
//   Cp  = the [P W] part of the active columns of Cx, orthogonalised against Cx and among
//         themselves by classical Gram-Schmidt with re-orthogonalisation; dependent columns dropped
//   coef (s, m + kept) row-major = [Cx | Cp].
// Replaces a host loop that cost ~7 ms per Rayleigh-Ritz step (plus two PCIe round trips).
__global__ void __launch_bounds__(1024) rr_coef_kernel(const double *__restrict__ evT, int s, int m,
                                                       const int *__restrict__ act, int q, double *__restrict__ coef,
                                                       int *__restrict__ kept_out) {
    extern __shared__ double sm[];
    double *cx = sm;                     // [m][s]
    double *qq = cx + (size_t)m * s;     // [q][s]
    double *dots = qq + (size_t)q * s;   // [m + q]
    __shared__ int s_kept;
    __shared__ double s_n0, s_n1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int i = tid; i < m * s; i += blockDim.x) cx[i] = evT[i];
    __syncthreads();
    for (int i = tid; i < q * s; i += blockDim.x) {
        const int j = i / s, r = i - j * s;
        qq[i] = r >= m ? cx[(size_t)act[j] * s + r] : 0.0;
    }
    if (tid == 0) s_kept = 0;
    __syncthreads();
    for (int j = 0; j < q; j++) {
        double *col = qq + (size_t)j * s;
        const int kept = s_kept;
        if (warp == 0) {
            double a = 0.0;
            for (int i = lane; i < s; i += 32) a += col[i] * col[i];
#pragma unroll
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s_n0 = sqrt(a);
        }
        __syncthreads();
        const int nb = m + kept;
        for (int rep = 0; rep < 2; rep++) {
            for (int b = warp; b < nb; b += nwarp) {
                const double *bv = b < m ? cx + (size_t)b * s : qq + (size_t)(b - m) * s;
                double a = 0.0;
                for (int i = lane; i < s; i += 32) a += bv[i] * col[i];
#pragma unroll
                for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0) dots[b] = a;
            }
            __syncthreads();
            if (tid < s) {
                double v = col[tid];
                for (int b = 0; b < m; b++) v -= dots[b] * cx[(size_t)b * s + tid];
                for (int b = 0; b < kept; b++) v -= dots[m + b] * qq[(size_t)b * s + tid];
                col[tid] = v;
            }
            __syncthreads();
        }
        if (warp == 0) {
            double a = 0.0;
            for (int i = lane; i < s; i += 32) a += col[i] * col[i];
#pragma unroll
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s_n1 = sqrt(a);
        }
        __syncthreads();
        const double n0 = s_n0, n1 = s_n1;
        const bool keep = n0 > 0.0 && !(n1 < 1e-8 * n0) && !(n1 < 1e-14);
        if (keep) {
            double *dst = qq + (size_t)kept * s;
            if (tid < s) dst[tid] = col[tid] / n1;  // kept <= j: never overwrites a column still to be processed
        }
        __syncthreads();
        if (tid == 0 && keep) s_kept = kept + 1;
        __syncthreads();
    }
    // an even number of P columns: the W block that follows [X P] in the basis then starts at a 16-byte
    // aligned column (vector loads of the SpMM, cp.async of the block update) and the coefficient matrix
    // has an even leading dimension; the dropped direction is the one of the last active column
    const int kept = s_kept & ~1, w = m + kept;
    for (int i = tid; i < s * w; i += blockDim.x) {
        const int r = i / w, cidx = i - r * w;
        coef[i] = cidx < m ? cx[(size_t)cidx * s + r] : qq[(size_t)(cidx - m) * s + r];
    }
    if (tid == 0) *kept_out = kept;
}

// G (s x s) from the carried [X P] block (w0 x w0) and the fresh block column Gw = S^T (A W) (s x mw):
// the W-W block is symmetrised (its two triangles come from different roundings of the same products)
__global__ void rr_gram_assemble(int s, int w0, const double *__restrict__ gxp, const double *__restrict__ gw,
                                 double *__restrict__ g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s * s) return;
    const int i = idx / s, j = idx - i * s, mw = s - w0;
    double v;
    if (i < w0 && j < w0) v = 0.5 * (gxp[i * w0 + j] + gxp[j * w0 + i]);
    else if (j >= w0 && i < w0) v = gw[i * mw + (j - w0)];
    else if (i >= w0 && j < w0) v = gw[j * mw + (i - w0)];
    else v = 0.5 * (gw[i * mw + (j - w0)] + gw[j * mw + (i - w0)]);
    g[idx] = v;
}

struct PhaseTimer {
    lb_ctx *c;
    double t0 = 0;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    explicit PhaseTimer(lb_ctx *ctx) : c(ctx) {}
    void start() {
        if (!c->trace) return;
        cudaStreamSynchronize(c->stream);
        t0 = wall_ms();
    }
    void stop(int slot) {
        if (!c->trace) return;
        cudaStreamSynchronize(c->stream);
        const double t = wall_ms();
        acc[slot] += t - t0;
        t0 = t;
    }
};

struct EigStats {
    int iterations = 0, converged = 0, levels = 0, block = 0;
    double residual = 0, setup_ms = 0, solve_ms = 0;
};

// One level of the (nested) iteration: LOBPCG on (A, B) with the multigrid cycle started at
// `lvl` as preconditioner.  x0 (n, m): optional initial block; x_out (n, m, ld = m): result block.
static EigStats lobpcg_core(lb_ctx *c, const lb_mat *A, const lb_mat *B, Amg *amg, int lvl, const double *x0, int ldx0,
                            int k, int m, double tol, int maxit, std::vector<double> &lam, double *x_out,
                            DistOps *D = nullptr, int64_t row0 = 0) {
    const int64_t n = A->n;
    EigStats st;
    const int ld = 3 * m;
    st.block = m;
    st.levels = (int)amg->levels.size();

    cudaEvent_t e0, e1;
    LB_CUDA(cudaEventCreate(&e0));
    LB_CUDA(cudaEventCreate(&e1));
    LB_CUDA(cudaEventRecord(e0, c->stream));

    // the eight big blocks live in the context's persistent workspace (common.cuh)
    const size_t blk = ((size_t)n * ld + 31) & ~(size_t)31, small = ((size_t)n * m + 31) & ~(size_t)31;
    struct Blk {
        double *p;
    };
    double *wsp = static_cast<double *>(ctx_workspace(c, (6 * blk + 2 * small) * sizeof(double)));
    Blk S[2] = {{wsp}, {wsp + blk}}, AS[2] = {{wsp + 2 * blk}, {wsp + 3 * blk}}, BS[2] = {{wsp + 4 * blk}, {wsp + 5 * blk}};
    Blk Rbuf{wsp + 6 * blk}, tmp{wsp + 6 * blk + small};
    DBuf<double> G(c, (size_t)ld * ld), evd(c, ld), lam_d(c, m), coef(c, (size_t)ld * 2 * m), dots(c, 2 * m);
    DBuf<int> idx_d(c, m), kept_d(c, 1);
    std::vector<double> hG, hC, hQ, coefh, rr(2 * m);
    lam.assign(m, 0.0);
    std::vector<int> act(m), idx(m);
    int cur = 0;

    phase(c, "lobpcg: work blocks allocated");
    // ---- initial block: constants + pseudo-random, B-orthonormalised, one Rayleigh-Ritz
    if (x0) {
        copy_cols(c, n, m, x0, ldx0, S[0].p, ld);  // prolonged coarse-level eigenvectors
    } else {
        fill_random(c, n, m, S[0].p, ld, 0x1234567ull, row0);
        LB_LAUNCH(c, set_column, cdiv(n, 256), 256, 0, n, S[0].p, ld, 0, 1.0);
    }
    int kept = b_orthonormalize(c, D, B, n, m, S[0].p, ld, BS[0].p, ld, tmp.p);
    LB_REQUIRE(kept == m, "initial block is rank deficient");
    phase(c, "lobpcg: initial block orthonormalised");
    d_spmm(c, D, A, S[0].p, ld, AS[0].p, ld, m);
    // [X P]^T A [X P] of the CURRENT basis, carried from the previous Rayleigh-Ritz step through the small
    // matrices (= coef^T G coef): the next Gram matrix then only needs its W block column from the big
    // blocks, S^T (A W), half the DMMA work of the symmetric s x s product (2.1 -> 1.1 ms at level 9).
    // Rounding differences to the directly computed blocks are O(eps ||G||) per step and do not compound
    // beyond a sum over the iterations (the W column and A [X P] itself are fresh every time).
    DBuf<double> Gkeep(c, (size_t)ld * ld), Gxp(c, (size_t)4 * m * m), Gw(c, (size_t)ld * m), Tsm(c, (size_t)ld * 2 * m);
    int gxp_w = 0;  // order of Gxp (0: not available)
    auto rayleigh_ritz = [&](int s, int mp_hint, const std::vector<int> &active_cols, int &mp_new) {
        // G = S^T A S (s x s); eigenvectors -> Cx; Cp from the active columns
        const int mw_blk = s - gxp_w;
        if (gxp_w > 0 && mw_blk > 0 && mw_blk <= m && s >= 2 * m) {
            d_gram(c, D, n, s, S[cur].p, ld, mw_blk, AS[cur].p + gxp_w, ld, Gw.p);  // (s x mw)
            LB_LAUNCH(c, rr_gram_assemble, cdiv(s * s, 256), 256, 0, s, gxp_w, Gxp.p, Gw.p, G.p);
        } else {
            d_gram(c, D, n, s, S[cur].p, ld, s, AS[cur].p, ld, G.p, true);
        }
        d2d(c, Gkeep.p, G.p, (size_t)s * s * sizeof(double));
        int info = sym_eig(c, s, G.p, evd.p);
        LB_REQUIRE(info == 0, "Rayleigh-Ritz eigen-decomposition failed (info=%d)", info);
        const int q = (int)active_cols.size();
        const bool want_p = s > m && q > 0;
        const size_t smem = ((size_t)(m + (want_p ? q : 0)) * s + m + q) * sizeof(double);
        int w;
        if (smem <= 200 * 1024) {
            // coefficient matrix built on the device; the host only needs lambda and the count
            if (want_p) h2d(c, idx_d.p, active_cols.data(), q * sizeof(int));
            LB_CUDA(cudaFuncSetAttribute(rr_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LB_LAUNCH(c, rr_coef_kernel, 1, 1024, smem, G.p, s, m, idx_d.p, want_p ? q : 0, coef.p, kept_d.p);
            d2h(c, lam.data(), evd.p, m * sizeof(double));
            int hk = 0;
            read_back(c, &hk, kept_d.p, 1);
            mp_new = hk;
            w = m + mp_new;
        } else {
            hG.resize((size_t)m * s);
            d2h(c, hG.data(), G.p, hG.size() * sizeof(double));  // rows 0..m-1 = the m smallest eigenvectors
            d2h(c, lam.data(), evd.p, m * sizeof(double));
            sync(c);
            hC.assign((size_t)s * m, 0.0);  // Cx (s x m) row-major
            for (int j = 0; j < m; j++)
                for (int i = 0; i < s; i++) hC[(size_t)i * m + j] = hG[(size_t)j * s + i];
            mp_new = 0;
            if (want_p) {
                hQ.assign((size_t)s * q, 0.0);
                for (int a = 0; a < q; a++)
                    for (int i = m; i < s; i++) hQ[(size_t)i * q + a] = hC[(size_t)i * m + active_cols[a]];
                mp_new = host_orth(s, m, hC, q, hQ) & ~1;  // even, like rr_coef_kernel
            }
            w = m + mp_new;
            coefh.assign((size_t)s * w, 0.0);
            for (int i = 0; i < s; i++) {
                for (int j = 0; j < m; j++) coefh[(size_t)i * w + j] = hC[(size_t)i * m + j];
                for (int j = 0; j < mp_new; j++) coefh[(size_t)i * w + m + j] = hQ[(size_t)i * q + j];
            }
            h2d(c, coef.p, coefh.data(), coefh.size() * sizeof(double));
        }
        {  // Gxp = coef^T (G coef), w x w: two small DMMA products
            ProfScope prof(c, PROF_TRSM, 4.0 * s * s * w, s, w);
            update_dmma(c, s, s, Gkeep.p, s, w, coef.p, w, 1.0, 0.0, Tsm.p, w);
            gram_dmma(c, s, w, coef.p, w, w, Tsm.p, w, Gxp.p, false);
            gxp_w = w;
        }
        const int nxt = cur ^ 1;
        update(c, n, s, S[cur].p, ld, w, coef.p, w, 1.0, 0.0, S[nxt].p, ld);
        // A [X P] and B [X P] are recomputed by SpMM (HBM-bound, ~2 ms at level 9) instead of
        // being carried through two more (n, s) x (s, w) products (~13 ms): cheaper and drift-free
        d_spmm(c, D, A, S[nxt].p, ld, AS[nxt].p, ld, w);
        d_spmm(c, D, B, S[nxt].p, ld, BS[nxt].p, ld, w);
        h2d(c, lam_d.p, lam.data(), m * sizeof(double));
        sync(c);  // coefh / lam are host buffers
        cur = nxt;
        (void)mp_hint;
    };
    int mp = 0;
    {
        std::vector<int> none;
        rayleigh_ritz(m, 0, none, mp);
    }
    phase(c, "lobpcg: first Rayleigh-Ritz");
    const bool f32_cycle = !D && amg_prepare_f32(*amg);
    phase(c, "lobpcg: single-precision hierarchy");

    double worst = 0.0;
    int nconv_k = 0;
    PhaseTimer pt(c);
    constexpr int ortho_passes = 2;  // block Gram-Schmidt against [X P]: "twice is enough" (one pass diverged, see below)
    for (int it = 0; it < maxit; it++) {
        st.iterations = it;
        pt.start();
        // ---- residual norms of all m columns
        residual_norms(c, n, m, lam_d.p, AS[cur].p, ld, BS[cur].p, ld, dots.p);
        if (D) d_allreduce(c, D, dots.p, 2 * m);
        read_back(c, rr.data(), dots.p, 2 * m);
        double lam_mean = 0;
        for (int j = 0; j < k; j++) lam_mean += std::fabs(lam[j]);
        lam_mean /= k;
        if (!(lam_mean > 0)) lam_mean = 1.0;
        worst = 0.0;
        nconv_k = 0;
        int ma = 0;
        for (int j = 0; j < m; j++) {
            const double scale = std::max(std::fabs(lam[j]), lam_mean);
            const double rel = std::sqrt(rr[j]) / (scale * std::sqrt(rr[m + j]));
            const bool conv = rel <= tol;
            act[j] = !conv;
            if (j < k) {
                nconv_k += conv;
                worst = std::max(worst, std::isfinite(rel) ? rel : INFINITY);
            }
            if (!conv) idx[ma++] = j;
        }
        // keep the active block even-sized: rows of the (n, ma) work blocks stay 16-byte aligned (bulk
        // copies / 16-byte loads of the SpMM); the extra column is a converged one, harmless to refine
        if ((ma & 1) && ma < m)
            for (int j = 0; j < m; j++)
                if (!act[j]) {
                    act[j] = 1;
                    idx[ma++] = j;
                    std::sort(idx.begin(), idx.begin() + ma);
                    break;
                }
        if (c->trace)
            fprintf(stderr, "[lb trace] lobpcg it %3d: max res(first k) %.3e, converged %d/%d, active %d, P %d\n", it,
                    worst, nconv_k, k, ma, mp);
        pt.stop(0);
        if (nconv_k == k || !std::isfinite(worst)) break;
        if (it == maxit - 1) break;
        // ---- W = precond(R_active), orthogonalised against [X P], B-orthonormalised
        std::vector<int> active_cols(idx.begin(), idx.begin() + ma);
        h2d(c, idx_d.p, idx.data(), ma * sizeof(int));
        const int w0 = m + mp;
        double *W = S[cur].p + w0, *AW = AS[cur].p + w0, *BW = BS[cur].p + w0;
        if (!D && f32_cycle) {
            // single-precision multigrid cycle: the residual block is converted as it is compacted, the
            // cycle's last kernel writes W in double (amg.cu)
            const int mpad = (ma + 3) & ~3;
            float *Rf = reinterpret_cast<float *>(Rbuf.p);
            residual_cols_f32(c, n, ma, idx_d.p, lam_d.p, AS[cur].p, ld, BS[cur].p, ld, Rf, mpad);
            amg_apply_f32(*amg, Rf, mpad, W, ld, mpad, ma, lvl);
        } else {
            residual_cols(c, n, ma, idx_d.p, lam_d.p, AS[cur].p, ld, BS[cur].p, ld, Rbuf.p, ma);
            if (D) dist_precond(c, D, Rbuf.p, ma, W, ld, ma);
            else amg_apply(*amg, Rbuf.p, ma, W, ld, ma, lvl);
        }
        pt.stop(1);
        // block Gram-Schmidt against [X P], twice ("twice is enough").  A single pass was measured to
        // lose orthogonality near convergence: the row-partitioned run (weaker block-Jacobi
        // preconditioner) diverged after ~50 iterations with one pass and converges with two.
        // (Round 2: making the second pass conditional on the Kahan-Parlett test - surviving norm
        // >= 0.7 of the column, from the first pass's coefficients and the diagonal of W^T B W - does not
        // pay: the preconditioned residuals lose more than that in 24 of 26 iterations at level 9.)
        for (int rep = 0; rep < (it < 2 ? 2 : ortho_passes); rep++) {
            d_gram(c, D, n, w0, BS[cur].p, ld, ma, W, ld, G.p);                   // (w0 x ma)
            update(c, n, w0, S[cur].p, ld, ma, G.p, ma, -1.0, 1.0, W, ld);        // W -= [X P] G
        }
        pt.stop(2);
        const int mw = b_orthonormalize(c, D, B, n, ma, W, ld, BW, ld, tmp.p, false);
        if (mw == 0) break;  // nothing left to add: stagnation
        pt.stop(3);
        d_spmm(c, D, A, W, ld, AW, ld, mw);
        pt.stop(4);
        // ---- Rayleigh-Ritz on [X P W]
        rayleigh_ritz(w0 + mw, mp, active_cols, mp);
        pt.stop(5);
    }
    if (c->trace)
        fprintf(stderr, "[lb trace] lobpcg phases (ms): residual %.1f | precond %.1f | ortho-vs-XP %.1f | B-orthonormalize %.1f | A*W %.1f | Rayleigh-Ritz+update %.1f\n",
                pt.acc[0], pt.acc[1], pt.acc[2], pt.acc[3], pt.acc[4], pt.acc[5]);
    phase(c, "lobpcg: iterations");
    st.iterations += 1;
    st.converged = nconv_k;
    st.residual = worst;

    LB_CUDA(cudaEventRecord(e1, c->stream));
    LB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    st.solve_ms = ms;

    copy_cols(c, n, m, S[cur].p, ld, x_out, m);
    return st;
}

// Driver: locality renumbering, AMG setup, nested iteration (coarse-level eigenvectors prolonged
// as the initial block of the next finer level - the eigen-analogue of full multigrid), output.
static EigStats lobpcg(lb_ctx *c, const lb_mat *A0, const lb_mat *B0, int k, double sigma, double tol, int maxit,
                       double *h_evals, double *h_evecs) {
    const int64_t n = A0->n;
    HostPrefault prefault(h_evecs, (size_t)n * k * sizeof(double));  // overlaps the whole solve
    phase(c, "(enter eigs)");
    // the solver iterates in the locality numbering the assembled matrices are stored in (Morton
    // order of the mesh: the SpMM gathers hit L1/L2 instead of DRAM, ncu: 4.0x -> 1.1x of the
    // algorithmic traffic); a user-assigned operand is permuted into it.  Results return in the
    // caller's order.
    MatView va, vb;
    std::shared_ptr<lb_order> ord = common_numbering(c, A0, B0, va, vb);
    const lb_mat *A = va.m, *B = vb.m;
    const bool reorder = ord != nullptr;
    int m = ((k + std::max(6, (k + 3) / 4) + 7) / 8) * 8;

    // preconditioner on K = A - sigma*B (SPD for sigma < 0)
    const double shift = sigma < 0 ? -sigma : 1e-2;
    AmgOptions opt;
    phase(c, "eigs: renumber A, B");
    auto amg = amg_setup(c, mat_axpby(c, A, 1.0, B, shift), m, opt);
    phase(c, "eigs: AMG setup");

    // ---- nested iteration: coarse pencils (K_l, B_l), B_{l+1} = R_l B_l P_l (Galerkin, like K_l).
    // K_l = A_l + shift*B_l has the eigenvectors of (A_l, B_l); only the vectors are carried up.
    const int nlev = (int)amg->levels.size();
    int depth = 0;  // number of coarse levels that get their own eigensolve
    // pays off only when the fine level dwarfs the per-iteration fixed cost (syevd, launches)
    while (n >= 1000000 && depth + 1 < nlev - 1 && amg->levels[depth + 1].K->n >= std::max<int64_t>(8 * m, 4000)) depth++;
    std::vector<std::unique_ptr<lb_mat>> Bl(depth + 1);
    for (int l = 0; l < depth; l++) {
        const lb_mat *bl = l == 0 ? B : Bl[l].get();
        auto BP = spgemm(c, bl, amg->levels[l].P.get());
        Bl[l + 1] = spgemm(c, amg->levels[l].R.get(), BP.get());
        Bl[l + 1]->ncols = -1;
    }
    phase(c, "eigs: coarse mass (Galerkin)");
    EigStats st, stc;
    std::vector<double> lam;
    DBuf<double> xc, xf;  // coarse result, prolonged initial block
    double coarse_ms = 0;
    for (int l = depth; l >= 1; l--) {
        const lb_mat *Kl = amg->levels[l].K.get();
        DBuf<double> xo(c, (size_t)Kl->n * m);
        // only a starting block, the fine level re-converges: 1e-2 is as good a start as 1e-3 (27 fine iterations
        // at level 9 either way) for 8 instead of 11 latency-bound coarse iterations per level
        stc = lobpcg_core(c, Kl, Bl[l].get(), amg.get(), l, xf.p, m, k, m, std::max(tol, 1e-2), 60, lam, xo.p);
        coarse_ms += stc.solve_ms;
        if (c->trace)
            fprintf(stderr, "[lb trace] nested level %d (n=%lld): %d iterations, residual %.2e, %.1f ms\n", l,
                    (long long)Kl->n, stc.iterations, stc.residual, stc.solve_ms);
        // prolong: X_{l-1} = P_{l-1} X_l
        const lb_mat *P = amg->levels[l - 1].P.get();
        xf.alloc(c, (size_t)P->n * m);
        spmm(c, P, xo.p, m, xf.p, m, m);
    }
    phase(c, "eigs: nested coarse solves");
    DBuf<double> xout(c, (size_t)n * m);
    st = lobpcg_core(c, A, B, amg.get(), 0, depth ? xf.p : nullptr, m, k, m, tol, maxit, lam, xout.p);
    phase(c, "eigs: fine-level LOBPCG");
    st.setup_ms = amg->setup_ms;
    st.solve_ms += coarse_ms;

    for (int j = 0; j < k; j++) h_evals[j] = lam[j];
    prefault.wait();
    if (!h_evecs) {  // eigenvalues only (batched ShapeDNA): no gather, no download
        sync(c);
        return st;
    }
    if (reorder) {
        // row i of the caller's numbering = row inv[i] of the renumbered block
        DBuf<double> out(c, (size_t)n * k);
        gather_rows(c, n, k, ord->inv.p, xout.p, m, out.p, k);
        d2h_large(c, h_evecs, out.p, (size_t)n * k * sizeof(double));
    } else {
        DBuf<double> out(c, (size_t)n * k);
        copy_cols(c, n, k, xout.p, m, out.p, k);
        d2h_large(c, h_evecs, out.p, (size_t)n * k * sizeof(double));
    }
    phase(c, "eigs: un-renumber + D2H");
    return st;
}

// ---- row-partitioned driver ------------------------------------------------------------------------
// Every rank holds the full (identical) A and B - assembled redundantly, 1-30 ms - renumbers them,
// keeps its contiguous row block, preconditions with an AMG hierarchy of its own diagonal block
// (block-Jacobi / non-overlapping additive Schwarz: no communication inside the cycle) and runs the
// same LOBPCG: identical small dense problems on every rank (all-reduced Gram matrices are bitwise
// equal), so no rank needs to be told what the others decided.
static EigStats lobpcg_dist(lb_ctx *c, const DistCtx *dist, const lb_mat *A0, const lb_mat *B0, int k, double sigma,
                            double tol, int maxit, double *h_evals, double *h_evecs) {
    const int64_t n = A0->n;
    HostPrefault prefault(h_evecs, (size_t)n * k * sizeof(double));
    MatView va, vb;
    std::shared_ptr<lb_order> ord = common_numbering(c, A0, B0, va, vb);
    const lb_mat *A = va.m, *B = vb.m;
    const bool reorder = ord != nullptr;
    const int m = ((k + std::max(6, (k + 3) / 4) + 7) / 8) * 8;
    const int world = dist->world, rank = dist->rank;
    const int64_t rpr = (n + world - 1) / world;
    const int64_t r0 = std::min(n, rank * rpr), r1 = std::min(n, r0 + rpr);
    LB_REQUIRE(r1 - r0 >= 4 * m, "row block of rank %d too small (%lld rows) for a block of %d", rank, (long long)(r1 - r0), m);
    const double shift = sigma < 0 ? -sigma : 1e-2;
    auto Ar = row_block(c, A, r0, r1, world * rpr);
    auto Br = row_block(c, B, r0, r1, world * rpr);  // general CSR with global columns (also when B is diagonal)
    // The hierarchy of the FULL operator on every rank (set-up is 10-20 ms, redundant like the
    // assembly), applied column-parallel (dist_precond): the iteration count equals the single-GPU
    // solver's (measured on 2 GPUs: 40 at level 9, 46 on the 121^3 cube; a per-rank block-Jacobi
    // hierarchy needed 189 / 75).  The nested-iteration start is computed redundantly on the small
    // coarse levels, so every rank starts from the same block.
    AmgOptions opt;
    std::unique_ptr<Amg> amg = amg_setup(c, mat_axpby(c, A, 1.0, B, shift), m, opt);
    DBuf<double> x_start;  // (n, m) prolonged coarse eigenvectors, identical on all ranks
    bool have_start = false;
    {
        const int nlev = (int)amg->levels.size();
        int depth = 0;
        while (n >= 1000000 && depth + 1 < nlev - 1 && amg->levels[depth + 1].K->n >= std::max<int64_t>(8 * m, 4000)) depth++;
        std::vector<std::unique_ptr<lb_mat>> Bl(depth + 1);
        for (int l = 0; l < depth; l++) {
            const lb_mat *bl = l == 0 ? B : Bl[l].get();
            auto BP = spgemm(c, bl, amg->levels[l].P.get());
            Bl[l + 1] = spgemm(c, amg->levels[l].R.get(), BP.get());
            Bl[l + 1]->ncols = -1;
        }
        std::vector<double> lamc;
        for (int l = depth; l >= 1; l--) {
            const lb_mat *Kl = amg->levels[l].K.get();
            DBuf<double> xo(c, (size_t)Kl->n * m);
            lobpcg_core(c, Kl, Bl[l].get(), amg.get(), l, have_start ? x_start.p : nullptr, m, k, m, std::max(tol, 1e-3), 60,
                        lamc, xo.p);
            const lb_mat *P = amg->levels[l - 1].P.get();
            x_start.alloc(c, (size_t)P->n * m);
            spmm(c, P, xo.p, m, x_start.p, m, m);
            have_start = true;
        }
    }
    DistOps D;
    D.d = dist;
    D.rpr = rpr;
    D.n_local = r1 - r0;
    D.r0 = r0;
    D.r1 = r1;
    D.wcap = 2 * m;
    D.pack.alloc((size_t)rpr * k);
    LB_CUDA(cudaMemsetAsync(D.pack.p, 0, D.pack.n * sizeof(double), c->stream));
    D.gath.alloc((size_t)world * rpr * k);
    D.red.alloc((size_t)9 * m * m + 4 * m);
    D.full_amg = amg.get();
    D.n_full = n;
    D.tpack.alloc((size_t)(r1 - r0) * m);
    D.tfull_r.alloc((size_t)n * ((m + world - 1) / world));
    D.tfull_z.alloc((size_t)n * ((m + world - 1) / world));
    if (c->trace) fprintf(stderr, "[lb trace] rank %d: rows [%lld, %lld) of %lld, AMG levels %zu (replicated, column-parallel)\n",
                          rank, (long long)r0, (long long)r1, (long long)n, amg->levels.size());
    std::vector<double> lam;
    DBuf<double> xloc(c, (size_t)(r1 - r0) * m);
    EigStats st = lobpcg_core(c, Ar.get(), Br.get(), amg.get(), 0, have_start ? x_start.p + (size_t)r0 * m : nullptr, m, k,
                              m, tol, maxit, lam, xloc.p, &D, r0);
    st.setup_ms = amg->setup_ms;
    for (int j = 0; j < k; j++) h_evals[j] = lam[j];
    if (!h_evecs) {
        sync(c);
        return st;
    }
    // all-gather the k eigenvector columns, undo the renumbering, return the full array on every rank
    DBuf<double> out(c, (size_t)n * k);
    copy_cols(c, r1 - r0, k, xloc.p, m, D.pack.p, k);
    dist_allgather(c, dist, D.pack.p, D.gath.p, (size_t)rpr * k);
    prefault.wait();
    if (reorder) gather_rows(c, n, k, ord->inv.p, D.gath.p, k, out.p, k);
    else d2d(c, out.p, D.gath.p, (size_t)n * k * sizeof(double));
    d2h_large(c, h_evecs, out.p, (size_t)n * k * sizeof(double));
    return st;
}

// ---- dense path for tiny problems (n too small for a 3m-wide basis) -----------------------------
__global__ void densify_sym(int64_t n, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                            const double *__restrict__ val, double *__restrict__ dense) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = ptr[i]; p < ptr[i + 1]; p++) dense[i * n + idx[p]] = val[p];
}

static void dense_eigs(lb_ctx *c, const lb_mat *A, const lb_mat *B, int k, double *h_evals, double *h_evecs) {
    const int n = (int)A->n;
    DBuf<double> dA(c, (size_t)n * n), dB(c, (size_t)n * n), w(c, n);
    dA.zero();
    dB.zero();
    LB_LAUNCH(c, densify_sym, cdiv(n, 128), 128, 0, (int64_t)n, A->indptr.p, A->indices.p, A->data.p, dA.p);
    LB_LAUNCH(c, densify_sym, cdiv(n, 128), 128, 0, (int64_t)n, B->indptr.p, B->indices.p, B->data.p, dB.p);
    if (!c->cusolver) {
        DBuf<double> dummy(c, 1), e(c, 1);
        dummy.zero();
        sym_eig(c, 1, dummy.p, e.p);  // creates the handle
    }
    cusolverDnHandle_t h = (cusolverDnHandle_t)c->cusolver;
    int lwork = 0;
    cusolverStatus_t s = cusolverDnDsygvd_bufferSize(h, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                                     CUBLAS_FILL_MODE_UPPER, n, dA.p, n, dB.p, n, w.p, &lwork);
    LB_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, "cusolver sygvd buffer query failed (%d)", (int)s);
    DBuf<double> work(c, lwork);
    DBuf<int> info(c, 1);
    s = cusolverDnDsygvd(h, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, dA.p, n, dB.p, n,
                         w.p, work.p, lwork, info.p);
    c->launches++;
    LB_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, "cusolver sygvd failed (%d)", (int)s);
    int hinfo = 0;
    read_back(c, &hinfo, info.p, 1);
    if (hinfo != 0) {
        set_error("dense generalized eigensolve failed (info=%d): mass matrix not positive definite?", hinfo);
        throw Error{LB_ERR_NOCONV};
    }
    d2h(c, h_evals, w.p, k * sizeof(double));
    if (!h_evecs) {
        sync(c);
        return;
    }
    // eigenvector j = column j column-major = row j row-major: transpose the first k rows into (n,k)
    std::vector<double> rows((size_t)k * n);
    d2h(c, rows.data(), dA.p, rows.size() * sizeof(double));
    sync(c);
    for (int j = 0; j < k; j++)
        for (int i = 0; i < n; i++) h_evecs[(size_t)i * k + j] = rows[(size_t)j * n + i];
}

}  // namespace lb

using namespace lb;

__global__ void max_abs_diff_kernel(int64_t count, const double *__restrict__ a, const double *__restrict__ b,
                                    unsigned long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d = i < count ? fabs(a[i] - b[i]) : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(d));
}

// self-check of the communicating primitives: row-partitioned SpMM / Gram against the same
// operations done redundantly on the full data of this rank.  errs[0] = max |dist SpMM - full SpMM|,
// errs[1] = max |dist Gram - full Gram| (relative to the largest entry).
extern "C" int lb_dist_selftest(lb_ctx *c, lb_mat *a, double *errs) {
    LB_API_BEGIN
    LB_REQUIRE(c && a && errs && c->dist, "lb_dist_selftest: needs a context with a communicator");
    DeviceGuard g(c->device);
    const int64_t n = a->n;
    const int w = 24, world = c->dist->world, rank = c->dist->rank;
    const int64_t rpr = (n + world - 1) / world, r0 = std::min(n, rank * rpr), r1 = std::min(n, r0 + rpr);
    auto Ar = row_block(c, a, r0, r1, world * rpr);
    DistOps D;
    D.d = c->dist;
    D.rpr = rpr;
    D.n_local = r1 - r0;
    D.r0 = r0;
    D.r1 = r1;
    D.wcap = w;
    D.red.alloc((size_t)w * w);
    DBuf<double> xf(c, (size_t)n * w), yf(c, (size_t)n * w), yd(c, (size_t)(r1 - r0) * w), gf(c, w * w), gd(c, w * w);
    DBuf<unsigned long long> mx(c, 2);
    mx.zero();
    fill_random(c, n, w, xf.p, w, 99, 0);
    spmm(c, a, xf.p, w, yf.p, w, w);
    d_spmm(c, &D, Ar.get(), xf.p + r0 * w, w, yd.p, w, w);
    LB_LAUNCH(c, max_abs_diff_kernel, cdiv((r1 - r0) * w, 256), 256, 0, (r1 - r0) * w, yd.p, yf.p + r0 * w, mx.p);
    gram(c, n, w, xf.p, w, w, yf.p, w, gf.p, false);
    d_gram(c, &D, r1 - r0, w, xf.p + r0 * w, w, w, yf.p + r0 * w, w, gd.p, false);
    LB_LAUNCH(c, max_abs_diff_kernel, cdiv(w * w, 256), 256, 0, (int64_t)w * w, gd.p, gf.p, mx.p + 1);
    unsigned long long h[2];
    read_back(c, h, mx.p, 2);
    std::memcpy(errs, h, 16);
    LB_API_END
}

extern "C" int lb_eigs(lb_ctx *c, lb_mat *a, lb_mat *b, int k, double sigma, double tol, int maxit, double *evals,
                       double *evecs, lb_info *info) {
    LB_API_BEGIN
    LB_REQUIRE(c && a && b && evals, "lb_eigs: NULL argument");  // evecs == NULL: eigenvalues only
    LB_REQUIRE(a->n == b->n, "stiffness and mass must have the same dimension");
    LB_REQUIRE(k >= 1 && k < a->n, "k must satisfy 1 <= k < n (n = %lld)", (long long)a->n);
    if (sigma > 0) {
        set_error("sigma > 0 (interior eigenvalues) is not supported by the block eigensolver; use sigma <= 0");
        return LB_ERR_UNSUPPORTED;
    }
    DeviceGuard g(c->device);
    if (tol <= 0) tol = 1e-9;
    if (maxit <= 0) maxit = 200;
    EigStats st;
    const int m = ((k + std::max(6, (k + 3) / 4) + 7) / 8) * 8;
    if (a->n <= std::max<int64_t>(4 * m, 600)) {
        // tiny problem: dense solve in the caller's numbering
        std::unique_ptr<lb_mat> pa, pb;
        if (a->permuted) pa = to_caller_order(c, a);
        if (b->permuted) pb = to_caller_order(c, b);
        dense_eigs(c, pa ? pa.get() : a, pb ? pb.get() : b, k, evals, evecs);
        st.converged = k;
    } else if (c->dist) {  // a communicator was attached (lb_comm_init): row-partitioned over its ranks
        st = lobpcg_dist(c, c->dist, a, b, k, sigma, tol, maxit, evals, evecs);
    } else {
        st = lobpcg(c, a, b, k, sigma, tol, maxit, evals, evecs);
    }
    if (info) {
        info->iterations = st.iterations;
        info->converged = st.converged;
        info->amg_levels = st.levels;
        info->reserved = st.block;
        info->residual = st.residual;
        info->setup_ms = st.setup_ms;
        info->solve_ms = st.solve_ms;
    }
    if (st.converged < k) {
        set_error("LOBPCG did not converge: %d of %d eigenpairs, max scaled residual %.3e after %d iterations",
                  st.converged, k, st.residual, st.iterations);
        return LB_ERR_NOCONV;
    }
    LB_API_END
}
