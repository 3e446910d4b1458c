// assembly.cu - linear-FEM stiffness / mass assembly on the device.
//
// Replaces lapy/solver.py:105-194 (_fem_tria), :196-308 (_fem_tria_aniso), :310-377
// (fem_tria_mass), :379-533 (_fem_tetra) and the SciPy COO->CSC conversion they end in.
//
// Everything below runs on the mesh's SOLVER LAYOUT (lb_mesh::v4m / t4m, built once at upload):
// vertices renumbered along the Morton curve, elements sorted by their smallest new vertex id.  Rows
// that are neighbours in memory are neighbours on the mesh, so the gathers of the element and row
// kernels (vertices, element records, incidence records) stay in L1 / L2 and the DRAM traffic is
// close to the compulsory figure; the matrices come out in the numbering the solvers iterate in
// (lb_mat::permuted) and are converted to the caller's canonical CSC only when downloaded.  The sum
// of every entry still runs over its addends in the reference's triplet order (ascending CALLER
// element id, then slot), so the values do not depend on the renumbering.
//
// Pipeline (all on the context's stream, one host read-back for nnz):
//   1. element pass      one thread per element: gather 3|4 vertices (one 32 B sector each),
//                        local cot / volume entries in the dtype of the caller's vertices with
//                        unfused IEEE ops, one 32 B (tria) / 96 B (tet) record per element,
//                        deterministic block partial sums of vol (for the degenerate clamp,
//                        solver.py:158-159) and the per-vertex incidence histogram.
//   2. incidence build   scan + fill + per-vertex sort: vertex -> (element*4 + corner), ascending.
//                        This IS the reference's triplet order restricted to one column.
//   3. row count         one thread per row: sorted unique neighbour keys in shared memory.
//   4. scan -> indptr; nnz read back; exact-size outputs allocated.
//   5. row fill          one thread per row accumulates (key, A, B) in COO input order into a
//                        shared-memory image of the block's contiguous CSR segment, which the
//                        block then streams out fully coalesced.  Every output byte is written
//                        exactly once; no atomics on values -> bit-reproducible.
// No sort of the 9T / 16T triplets is ever materialised (SURVEY.md §7 "Hard parts").
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace lb {

constexpr int kRowThreads = 128;  // rows per block in the row kernels

// ---- upload / conversion ----------------------------------------------------------------
template <class TIn>
__global__ void convert_vertices(const TIn *__restrict__ raw, int64_t nv, D4 *__restrict__ v4,
                                 float4 *__restrict__ v4f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    TIn x = raw[3 * i], y = raw[3 * i + 1], z = raw[3 * i + 2];
    st_d4(v4 + i, (double)x, (double)y, (double)z, 0.0);
    if (v4f) v4f[i] = make_float4((float)x, (float)y, (float)z, 0.f);
}

template <class TIn>
__global__ void convert_elements(const TIn *__restrict__ raw, int64_t nt, int k, int4 *__restrict__ t4,
                                 long long *__restrict__ minmax) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    if (e < nt) {
        long long a = raw[k * e], b = raw[k * e + 1], c = raw[k * e + 2];
        long long d = k == 4 ? (long long)raw[k * e + 3] : -1;
        lo = min(a, min(b, c));
        hi = max(a, max(b, c));
        if (k == 4) {
            lo = min(lo, d);
            hi = max(hi, d);
        }
        t4[e] = make_int4((int)a, (int)b, (int)c, (int)d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo != LLONG_MAX) {
        atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

// ---- deterministic block sum ------------------------------------------------------------
__device__ __forceinline__ double block_sum(double x) {
    __shared__ double ws[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = 0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < ((blockDim.x + 31) >> 5) ? ws[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    return s;  // valid in thread 0
}

// ---- element records --------------------------------------------------------------------
// triangle record (D4): a12, a23, a31, bii        (final fp64 values)
//   degenerate (vol < eps before the clamp): d12, d23, d31 numerators, w = -1
// tet record (3 x D4): a12 a13 a14 a23 | a24 a34 a11 a22 | a33 a44 bii lump   (A already /6)
//   degenerate: the six numerators in slots 0..5, bii = -1
enum { MODE_FEM = 0, MODE_ANISO = 1, MODE_MASS = 2 };

struct ElemConsts {
    double vol_mean;  // clamp value in the element dtype, widened
    double bii_deg;   // bii of a clamped element
    double lump_deg;  // lumped mass contribution of a clamped element (tets)
};

// local entries of one triangle (solver.py:145-169, :259-280, :342-358) in the dtype of the vertices:
// q12, q23, q31 (divided by vol unless degenerate), bii (-1 marks a degenerate element), vol
template <class T, int MODE>
__device__ __forceinline__ void tria_local(const Vec3<T> &p1, const Vec3<T> &p2, const Vec3<T> &p3, int64_t eo,
                                           const double *__restrict__ u1, const double *__restrict__ u2,
                                           const double *__restrict__ am, double &q12, double &q23, double &q31,
                                           double &bii, double &vol_d) {
    using E = Ex<T>;
    using ED = Ex<double>;
    Vec3<T> ec = vsub(p2, p1), ea = vsub(p3, p2), eb = vsub(p1, p3);  // v2mv1, v3mv2, v1mv3
    Vec3<T> cr = vcross(ea, eb);
    T s = E::sqrt(vdot(cr, cr));
    T vol = MODE == MODE_MASS ? E::mul((T)0.5, s) : E::mul((T)2, s);
    vol_d = (double)vol;
    const bool degen = vol < E::eps();
    if (MODE == MODE_FEM) {
        T d12 = vdot(ea, eb), d23 = vdot(eb, ec), d31 = vdot(ec, ea);
        if (!degen) {
            d12 = E::div(d12, vol);
            d23 = E::div(d23, vol);
            d31 = E::div(d31, vol);
        }
        q12 = (double)d12; q23 = (double)d23; q31 = (double)d31;
    } else if (MODE == MODE_ANISO) {
        // projections and the weighted dot run in fp64 (u1, u2, aniso_mat are fp64 arrays of the caller,
        // indexed by the CALLER's element id eo)
        Vec3<double> a = vwiden(ea), b = vwiden(eb), c = vwiden(ec);
        Vec3<double> w1 = {u1[3 * eo], u1[3 * eo + 1], u1[3 * eo + 2]};
        Vec3<double> w2 = {u2[3 * eo], u2[3 * eo + 1], u2[3 * eo + 2]};
        double m0 = am[2 * eo], m1 = am[2 * eo + 1];
        double a0 = vdot(w1, a), a1 = vdot(w2, a), b0 = vdot(w1, b), b1 = vdot(w2, b);
        double c0 = vdot(w1, c), c1 = vdot(w2, c);
        auto adot = [&](double x0, double x1, double y0, double y1) {
            return ED::add(ED::mul(ED::mul(x0, m0), y0), ED::mul(ED::mul(x1, m1), y1));
        };
        q12 = adot(a0, a1, b0, b1);
        q23 = adot(b0, b1, c0, c1);
        q31 = adot(c0, c1, a0, a1);
        if (!degen) {
            q12 = ED::div(q12, vol_d);
            q23 = ED::div(q23, vol_d);
            q31 = ED::div(q31, vol_d);
        }
    } else {
        q12 = q23 = q31 = 0.0;
    }
    bii = degen ? -1.0 : (double)E::div(vol, MODE == MODE_MASS ? (T)6 : (T)24);
}

template <class T, int MODE>
__global__ void __launch_bounds__(256) tria_element_kernel(
    const typename Ex<T>::V4 *__restrict__ v4, const int4 *__restrict__ t4, int64_t nt,
    const double *__restrict__ u1, const double *__restrict__ u2, const double *__restrict__ am,
    D4 *__restrict__ rec, int32_t *__restrict__ deg, double *__restrict__ partial) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double vol_d = 0.0;
    if (e < nt) {
        int4 ti = __ldg(t4 + e);
        Vec3<T> p1 = load_vertex<T>(v4, ti.x), p2 = load_vertex<T>(v4, ti.y), p3 = load_vertex<T>(v4, ti.z);
        double q12, q23, q31, bii;
        tria_local<T, MODE>(p1, p2, p3, ti.w, u1, u2, am, q12, q23, q31, bii, vol_d);
        st_d4(rec + e, q12, q23, q31, bii);
        if (deg) {
            atomicAdd(deg + ti.x, 1);
            atomicAdd(deg + ti.y, 1);
            atomicAdd(deg + ti.z, 1);
        }
    }
    double s = block_sum(vol_d);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

template <class T>
__global__ void __launch_bounds__(256) tet_element_kernel(
    const typename Ex<T>::V4 *__restrict__ v4, const int4 *__restrict__ t4, int64_t nt,
    D4 *__restrict__ rec, int32_t *__restrict__ deg, double *__restrict__ partial) {
    using E = Ex<T>;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double vol_d = 0.0;
    if (e < nt) {
        int4 ti = __ldg(t4 + e);
        Vec3<T> p1 = load_vertex<T>(v4, ti.x), p2 = load_vertex<T>(v4, ti.y);
        Vec3<T> p3 = load_vertex<T>(v4, ti.z), p4 = load_vertex<T>(v4, ti.w);
        Vec3<T> e1 = vsub(p2, p1), e2 = vsub(p3, p2), e3 = vsub(p1, p3);
        Vec3<T> e4 = vsub(p4, p1), e5 = vsub(p4, p2), e6 = vsub(p4, p3);
        T vol = fabs(vdot(e4, vcross(e1, e3)));
        vol_d = (double)vol;
        bool degen = vol < E::eps();
        T e11 = vdot(e1, e1), e22 = vdot(e2, e2), e33 = vdot(e3, e3);
        T e44 = vdot(e4, e4), e55 = vdot(e5, e5), e66 = vdot(e6, e6);
        T e12 = vdot(e1, e2), e13 = vdot(e1, e3), e14 = vdot(e1, e4), e15 = vdot(e1, e5);
        T e23 = vdot(e2, e3), e25 = vdot(e2, e5), e26 = vdot(e2, e6);
        T e34 = vdot(e3, e4), e36 = vdot(e3, e6);
        T a12 = E::add(E::mul(-e36, e26), E::mul(e23, e66));
        T a13 = E::add(E::mul(-e15, e25), E::mul(e12, e55));
        T a14 = E::sub(E::mul(e23, e26), E::mul(e36, e22));
        T a23 = E::add(E::mul(-e14, e34), E::mul(e13, e44));
        T a24 = E::sub(E::mul(e13, e34), E::mul(e14, e33));
        T a34 = E::add(E::mul(-e14, e13), E::mul(e11, e34));
        D4 *r = rec + 3 * e;
        if (degen) {
            st_d4(r, (double)a12, (double)a13, (double)a14, (double)a23);
            st_d4(r + 1, (double)a24, (double)a34, 0.0, 0.0);
            st_d4(r + 2, 0.0, 0.0, -1.0, 0.0);
        } else {
            a12 = E::div(a12, vol); a13 = E::div(a13, vol); a14 = E::div(a14, vol);
            a23 = E::div(a23, vol); a24 = E::div(a24, vol); a34 = E::div(a34, vol);
            T a11 = E::sub(E::sub(-a12, a13), a14);
            T a22 = E::sub(E::sub(-a12, a23), a24);
            T a33 = E::sub(E::sub(-a13, a23), a34);
            T a44 = E::sub(E::sub(-a14, a24), a34);
            const T six = (T)6;
            st_d4(r, (double)E::div(a12, six), (double)E::div(a13, six), (double)E::div(a14, six),
                  (double)E::div(a23, six));
            st_d4(r + 1, (double)E::div(a24, six), (double)E::div(a34, six), (double)E::div(a11, six),
                  (double)E::div(a22, six));
            st_d4(r + 2, (double)E::div(a33, six), (double)E::div(a44, six), (double)E::div(vol, (T)60),
                  (double)E::div(vol, (T)24));
        }
        if (deg) {
            atomicAdd(deg + ti.x, 1);
            atomicAdd(deg + ti.y, 1);
            atomicAdd(deg + ti.z, 1);
            atomicAdd(deg + ti.w, 1);
        }
    }
    double s = block_sum(vol_d);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// vol_mean = max(1e-4 * mean(vol), eps) in the element dtype (solver.py:158, :357, :436)
template <class T>
__global__ void finalize_consts(const double *__restrict__ partial, int nblocks, int64_t nt, int kind,
                                ElemConsts *__restrict__ out) {
    using E = Ex<T>;
    double s = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[i];
    s = block_sum(s);
    if (threadIdx.x == 0) {
        T mean = (T)(s / (double)nt);
        T vm = E::mul((T)0.0001, mean);
        if (!((double)vm > kEps)) vm = (T)kEps;  // python max(a, eps): a if a > eps else eps
        out->vol_mean = (double)vm;
        if (kind == LB_FEM_TETRA) {
            out->bii_deg = (double)E::div(vm, (T)60);
            out->lump_deg = (double)E::div(vm, (T)24);
        } else {
            out->bii_deg = (double)E::div(vm, kind == LB_FEM_TRIA_MASS ? (T)6 : (T)24);
            out->lump_deg = 2.0 * out->bii_deg;
        }
    }
}

// ---- incidence ----------------------------------------------------------------------------
__global__ void incidence_fill(const int4 *__restrict__ t4, int64_t nt, int k,
                               const int32_t *__restrict__ inc_ptr, int32_t *__restrict__ cursor,
                               int32_t *__restrict__ inc) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    int4 ti = __ldg(t4 + e);
    int code = (int)e * 4;
    inc[inc_ptr[ti.x] + atomicAdd(cursor + ti.x, 1)] = code;
    inc[inc_ptr[ti.y] + atomicAdd(cursor + ti.y, 1)] = code + 1;
    inc[inc_ptr[ti.z] + atomicAdd(cursor + ti.z, 1)] = code + 2;
    if (k == 4) inc[inc_ptr[ti.w] + atomicAdd(cursor + ti.w, 1)] = code + 3;
}

// assembly form of the incidence: one 16-byte record per (vertex, incident element) that already
// carries the other vertices of the element in the reference's triplet order for that column, so
// the row kernels never gather the element array:  {element*4 + corner, k0, k1, k2}
//   triangle corner 0: (t2, t3)   1: (t1, t3)   2: (t2, t1)          (solver.py:171-175)
//   tet      corner 0: (t2,t3,t4) 1: (t1,t3,t4) 2: (t2,t1,t4) 3: (t1,t2,t3)   (solver.py:472-497)
// `element` is the position in the sorted element array t4m; the reference's triplet order is the
// CALLER's element order: triangles carry that id in .w, tets look it up in eorig[] (inc_key)
// The records of one vertex are placed by an atomic cursor.  ncu (round 2): with one global atomic
// per incidence the kernel is latency bound (7 % issue utilisation, 113 us at level 9).  The elements
// are sorted by their smallest (new) vertex id, so the vertices of a chunk of consecutive elements
// lie in a narrow id window above the chunk's first key: ranks inside the chunk come from
// shared-memory atomics and only ONE global atomic per (chunk, vertex) reserves the chunk's range of
// the vertex's list; vertices outside the window fall back to a global atomic per incidence.  The
// order inside a list is arbitrary either way - the row kernels sort by the caller's element id.
constexpr int kIncWin = 2048;  // vertex window per chunk
constexpr int kIncEpt = 4;     // elements per thread (chunk = 256 * 4 elements)

__global__ void __launch_bounds__(256) incidence_fill4(const int4 *__restrict__ t4, int64_t nt, int k,
                                                        const int32_t *__restrict__ inc_ptr,
                                                        int32_t *__restrict__ cursor, int4 *__restrict__ inc4) {
    __shared__ int s_cnt[kIncWin];
    __shared__ int s_base[kIncWin];
    __shared__ int s_vbase;
    const int64_t e0 = (int64_t)blockIdx.x * (256 * kIncEpt);
    for (int i = threadIdx.x; i < kIncWin; i += 256) s_cnt[i] = 0;
    if (threadIdx.x == 0) {
        const int4 f = __ldg(t4 + e0);  // e0 < nt by the grid size
        int m = min(f.x, min(f.y, f.z));
        if (k == 4) m = min(m, f.w);
        s_vbase = m;
    }
    __syncthreads();
    const int vbase = s_vbase;
    int4 ti[kIncEpt];
    int rk[kIncEpt][4];
#pragma unroll
    for (int j = 0; j < kIncEpt; j++) {
        const int64_t e = e0 + (int64_t)j * 256 + threadIdx.x;
        ti[j] = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int cn = 0; cn < 4; cn++) rk[j][cn] = -1;
        if (e < nt) {
            ti[j] = __ldg(t4 + e);
            const int vs[4] = {ti[j].x, ti[j].y, ti[j].z, ti[j].w};
#pragma unroll
            for (int cn = 0; cn < 4; cn++) {
                if (cn < k) {
                    const int lv = vs[cn] - vbase;
                    if (lv >= 0 && lv < kIncWin) rk[j][cn] = atomicAdd(&s_cnt[lv], 1);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kIncWin; i += 256) {
        const int cnt = s_cnt[i];
        if (cnt) s_base[i] = atomicAdd(cursor + vbase + i, cnt);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kIncEpt; j++) {
        const int64_t e = e0 + (int64_t)j * 256 + threadIdx.x;
        if (e >= nt) continue;
        const int code = (int)e * 4;
        const int vs[4] = {ti[j].x, ti[j].y, ti[j].z, ti[j].w};
        int pos[4];
#pragma unroll
        for (int cn = 0; cn < 4; cn++) {
            pos[cn] = 0;
            if (cn < k) {
                const int v = vs[cn];
                pos[cn] = inc_ptr[v] + (rk[j][cn] >= 0 ? s_base[v - vbase] + rk[j][cn] : atomicAdd(cursor + v, 1));
            }
        }
        if (k == 3) {
            inc4[pos[0]] = make_int4(code, ti[j].y, ti[j].z, ti[j].w);
            inc4[pos[1]] = make_int4(code + 1, ti[j].x, ti[j].z, ti[j].w);
            inc4[pos[2]] = make_int4(code + 2, ti[j].y, ti[j].x, ti[j].w);
        } else {
            inc4[pos[0]] = make_int4(code, ti[j].y, ti[j].z, ti[j].w);
            inc4[pos[1]] = make_int4(code + 1, ti[j].x, ti[j].z, ti[j].w);
            inc4[pos[2]] = make_int4(code + 2, ti[j].y, ti[j].x, ti[j].w);
            inc4[pos[3]] = make_int4(code + 3, ti[j].x, ti[j].y, ti[j].z);
        }
    }
}

// position of an incidence record in the reference's triplet order of its column: caller's element
// id, then corner.  Triangles: .w is the caller's element id (one corner per element and vertex
// unless the element is degenerate: then the corner in .x breaks the tie - both fit 62 bits)
template <int K>
__device__ __forceinline__ long long inc_key(const int4 &q, const int32_t *__restrict__ eorig) {
    if (K == 3) return ((long long)q.w << 2) | (q.x & 3);
    return ((long long)__ldg(eorig + (q.x >> 2)) << 2) | (q.x & 3);
}

// per-vertex insertion sort of the (atomically ordered) incidence codes -> deterministic
__global__ void incidence_sort(const int32_t *__restrict__ inc_ptr, int32_t *__restrict__ inc, int64_t n) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int beg = inc_ptr[r], end = inc_ptr[r + 1];
    for (int i = beg + 1; i < end; i++) {
        int key = inc[i];
        int j = i - 1;
        while (j >= beg && inc[j] > key) {
            inc[j + 1] = inc[j];
            j--;
        }
        inc[j + 1] = key;
    }
}

// ---- sorted small-set helpers (thread-private slices of shared or global memory) ----------
__device__ __forceinline__ int find_slot(const int32_t *keys, int cnt, int key) {
    // rows hold ~7 (tria) / ~15 (tet) keys: a backwards linear scan beats a binary search
    int pos = cnt;
    while (pos > 0 && keys[pos - 1] >= key) pos--;
    return pos;  // first position with keys[pos] >= key
}

__device__ __forceinline__ void insert_key(int32_t *keys, int &cnt, int key) {
    int pos = find_slot(keys, cnt, key);
    if (pos < cnt && keys[pos] == key) return;
    for (int q = cnt; q > pos; q--) keys[q] = keys[q - 1];
    keys[pos] = key;
    cnt++;
}

__device__ __forceinline__ void accumulate(int32_t *keys, double *av, double *bv, int &cnt, int key,
                                           double a, double b, bool want_a, bool want_b) {
    int pos = find_slot(keys, cnt, key);
    if (pos < cnt && keys[pos] == key) {
        if (want_a) av[pos] = __dadd_rn(av[pos], a);
        if (want_b) bv[pos] = __dadd_rn(bv[pos], b);
        return;
    }
    for (int q = cnt; q > pos; q--) {
        keys[q] = keys[q - 1];
        if (want_a) av[q] = av[q - 1];
        if (want_b) bv[q] = bv[q - 1];
    }
    keys[pos] = key;
    if (want_a) av[pos] = a;  // first addend itself, like csr_sum_duplicates (keeps -0.0)
    if (want_b) bv[pos] = b;
    cnt++;
}

// ---- row count ----------------------------------------------------------------------------
// dynamic smem: cap int32 keys.  Blocks whose upper bound exceeds cap use the global scratch.
template <int K>
__global__ void __launch_bounds__(kRowThreads) row_count_kernel(
    const int32_t *__restrict__ inc_ptr, const int4 *__restrict__ inc4, int64_t n, int cap,
    int32_t *__restrict__ scratch, int32_t *__restrict__ row_nnz, int32_t *__restrict__ row_has,
    const int32_t *__restrict__ only_rows) {
    extern __shared__ int32_t skeys[];
    __shared__ int s_total;
    int64_t r = (int64_t)blockIdx.x * kRowThreads + threadIdx.x;
    int beg = 0, end = 0;
    // only_rows != NULL: second pass behind the fused kernel, only the rows it flagged (global scratch)
    const bool skip = r >= n || (only_rows && !only_rows[r]);
    if (!skip) {
        beg = inc_ptr[r];
        end = inc_ptr[r + 1];
    }
    int ninc = end - beg;
    int ub = ninc ? ninc * (K - 1) + 1 : 0;
    // block exclusive scan of ub
    __shared__ int wsum[kRowThreads / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = ub;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; w++) base += wsum[w];
    if (threadIdx.x == kRowThreads - 1) s_total = base + incl;
    __syncthreads();
    int off = base + incl - ub;
    int32_t *keys = (s_total <= cap && !only_rows) ? skeys + off : scratch + ((int64_t)beg * (K - 1) + r);
    int cnt = 0;
    if (ninc) insert_key(keys, cnt, (int)r);
    for (int p = beg; p < end; p++) {  // streaming: the records of a row are contiguous
        const int4 q = __ldg(inc4 + p);
        insert_key(keys, cnt, q.y);
        insert_key(keys, cnt, q.z);
        if (K == 4) insert_key(keys, cnt, q.w);
    }
    if (!skip) {
        row_nnz[r] = cnt;
        row_has[r] = ninc > 0;
    }
}

// ---- row fill -----------------------------------------------------------------------------
struct RowOut {
    const int32_t *indptr;
    int32_t *a_idx;
    double *a_val;  // may be NULL (mass only)
    int32_t *b_idx;
    double *b_val;  // may be NULL (lumped)
    const int32_t *lump_ptr;  // lumped mass: indptr of the diagonal matrix, or NULL
    int32_t *lump_idx;
    double *lump_val;
};

template <int K>
__global__ void __launch_bounds__(kRowThreads) row_fill_kernel(
    const D4 *__restrict__ rec, const int32_t *__restrict__ inc_ptr, int4 *inc4, const int32_t *__restrict__ eorig,
    int64_t n, int cap, const ElemConsts *__restrict__ consts, int degen_div_f32, RowOut out,
    const int32_t *__restrict__ only_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_a = reinterpret_cast<double *>(smem_raw);
    double *s_b = s_a + cap;
    int32_t *s_k = reinterpret_cast<int32_t *>(s_b + cap);

    const int64_t r0 = (int64_t)blockIdx.x * kRowThreads;
    const int64_t r = r0 + threadIdx.x;
    const int64_t rlast = min(r0 + (int64_t)kRowThreads, n);
    const int blk_beg = out.indptr[r0], blk_end = out.indptr[rlast];
    const int blk_nnz = blk_end - blk_beg;
    const bool want_a = out.a_val != nullptr, want_b = out.b_val != nullptr;
    const bool want_pat = want_a || want_b;
    // only_rows != NULL: second pass behind the fused kernels, only the flagged rows, straight to global
    const bool use_smem = want_pat && blk_nnz <= cap && !only_rows;

    if (r < n && (!only_rows || only_rows[r])) {
        const int rbeg = out.indptr[r];
        int32_t *keys;
        double *av, *bv;
        if (use_smem) {
            keys = s_k + (rbeg - blk_beg);
            av = s_a + (rbeg - blk_beg);
            bv = s_b + (rbeg - blk_beg);
        } else {  // oversized block (very high valence): accumulate straight into the output
            keys = (want_a ? out.a_idx : out.b_idx) + rbeg;
            av = out.a_val + rbeg;
            bv = out.b_val + rbeg;
        }
        int cnt = 0;
        double lump = 0.0;
        const int beg = inc_ptr[r], end = inc_ptr[r + 1];
        const int ninc = end - beg;
        // The incidence records were placed by atomics: restore the reference's triplet order
        // (CALLER element id ascending, inc_key).  Triangles with valence <= 8 are sorted in registers by a 19-comparator
        // network and their element records gathered together (8 independent loads in flight);
        // longer rows / tets sort their own segment in place and stream it.
        constexpr int RS = 8;
        int4 qi[RS];
        const bool in_regs = K == 3 && ninc <= RS;
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < RS; u++) qi[u] = u < ninc ? inc4[beg + u] : make_int4(INT_MAX, 0, 0, INT_MAX);
#define LB_CSWAP(i, j)                                                              \
    if (qi[i].w > qi[j].w || (qi[i].w == qi[j].w && qi[i].x > qi[j].x)) {           \
        const int4 tmp_ = qi[i];                                                    \
        qi[i] = qi[j];                                                              \
        qi[j] = tmp_;                                                               \
    }
            LB_CSWAP(0, 1) LB_CSWAP(2, 3) LB_CSWAP(4, 5) LB_CSWAP(6, 7) LB_CSWAP(0, 2) LB_CSWAP(1, 3) LB_CSWAP(4, 6)
            LB_CSWAP(5, 7) LB_CSWAP(1, 2) LB_CSWAP(5, 6) LB_CSWAP(0, 4) LB_CSWAP(3, 7) LB_CSWAP(1, 5) LB_CSWAP(2, 6)
            LB_CSWAP(1, 4) LB_CSWAP(3, 6) LB_CSWAP(2, 4) LB_CSWAP(3, 5) LB_CSWAP(3, 4)
#undef LB_CSWAP
        } else {
            for (int i = beg + 1; i < end; i++) {
                const int4 key = inc4[i];
                const long long kk = inc_key<K>(key, eorig);
                int j = i - 1;
                while (j >= beg && inc_key<K>(inc4[j], eorig) > kk) {
                    inc4[j + 1] = inc4[j];
                    j--;
                }
                inc4[j + 1] = key;
            }
        }
        auto body = [&](const int4 r4, const D4 q) {
            const int code = r4.x;
            const int e = code >> 2, c = code & 3;
            if (K == 3) {
                double a12 = q.x, a23 = q.y, a31 = q.z, bii = q.w;
                if (bii < 0.0) {  // clamped element (solver.py:159): divide by the global mean
                    const double vm = consts->vol_mean;
                    if (degen_div_f32) {
                        a12 = (double)__fdiv_rn((float)a12, (float)vm);
                        a23 = (double)__fdiv_rn((float)a23, (float)vm);
                        a31 = (double)__fdiv_rn((float)a31, (float)vm);
                    } else {
                        a12 = __ddiv_rn(a12, vm);
                        a23 = __ddiv_rn(a23, vm);
                        a31 = __ddiv_rn(a31, vm);
                    }
                    bii = consts->bii_deg;
                }
                const double bij = 0.5 * bii;
                lump += 2.0 * bii;  // vol/12 (vol/3 for fem_tria_mass) == 2*bii exactly
                int k0, k1;
                double x0, x1, xd;
                // column = corner c; rows in the reference's triplet order (solver.py:171-175)
                // the diagonal (row sum = 0, solver.py:167-169) is formed in the element dtype
                double m0, m1;
                k0 = r4.y;  // neighbours already in the triplet order of this column
                k1 = r4.z;
                if (c == 0) {
                    x0 = a12; x1 = a31; m0 = a12; m1 = a31;
                } else if (c == 1) {
                    x0 = a12; x1 = a23; m0 = a12; m1 = a23;
                } else {
                    x0 = a23; x1 = a31; m0 = a31; m1 = a23;
                }
                xd = degen_div_f32 ? (double)__fsub_rn(-(float)m0, (float)m1) : __dsub_rn(-m0, m1);
                if (want_pat) {
                    accumulate(keys, av, bv, cnt, k0, x0, bij, want_a, want_b);
                    accumulate(keys, av, bv, cnt, k1, x1, bij, want_a, want_b);
                    accumulate(keys, av, bv, cnt, (int)r, xd, bii, want_a, want_b);
                }
            } else {
                const D4 *rp = rec + 3 * (int64_t)e;
                D4 q0 = ldg_d4(rp), q1 = ldg_d4(rp + 1), q2 = ldg_d4(rp + 2);
                double a12 = q0.x, a13 = q0.y, a14 = q0.z, a23 = q0.w, a24 = q1.x, a34 = q1.y;
                double a11 = q1.z, a22 = q1.w, a33 = q2.x, a44 = q2.y, bii = q2.z, lmp = q2.w;
                if (bii < 0.0) {
                    const double vm = consts->vol_mean;
                    if (degen_div_f32) {
                        const float vf = (float)vm;
                        float f12 = __fdiv_rn((float)a12, vf), f13 = __fdiv_rn((float)a13, vf);
                        float f14 = __fdiv_rn((float)a14, vf), f23 = __fdiv_rn((float)a23, vf);
                        float f24 = __fdiv_rn((float)a24, vf), f34 = __fdiv_rn((float)a34, vf);
                        a11 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f12, f13), f14), 6.f);
                        a22 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f12, f23), f24), 6.f);
                        a33 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f13, f23), f34), 6.f);
                        a44 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f14, f24), f34), 6.f);
                        a12 = (double)__fdiv_rn(f12, 6.f); a13 = (double)__fdiv_rn(f13, 6.f);
                        a14 = (double)__fdiv_rn(f14, 6.f); a23 = (double)__fdiv_rn(f23, 6.f);
                        a24 = (double)__fdiv_rn(f24, 6.f); a34 = (double)__fdiv_rn(f34, 6.f);
                    } else {
                        double f12 = __ddiv_rn(a12, vm), f13 = __ddiv_rn(a13, vm), f14 = __ddiv_rn(a14, vm);
                        double f23 = __ddiv_rn(a23, vm), f24 = __ddiv_rn(a24, vm), f34 = __ddiv_rn(a34, vm);
                        a11 = __ddiv_rn(__dsub_rn(__dsub_rn(-f12, f13), f14), 6.0);
                        a22 = __ddiv_rn(__dsub_rn(__dsub_rn(-f12, f23), f24), 6.0);
                        a33 = __ddiv_rn(__dsub_rn(__dsub_rn(-f13, f23), f34), 6.0);
                        a44 = __ddiv_rn(__dsub_rn(__dsub_rn(-f14, f24), f34), 6.0);
                        a12 = __ddiv_rn(f12, 6.0); a13 = __ddiv_rn(f13, 6.0); a14 = __ddiv_rn(f14, 6.0);
                        a23 = __ddiv_rn(f23, 6.0); a24 = __ddiv_rn(f24, 6.0); a34 = __ddiv_rn(f34, 6.0);
                    }
                    bii = consts->bii_deg;
                    lmp = consts->lump_deg;
                }
                const double bij = 0.5 * bii;
                lump += lmp;
                int k0, k1, k2;
                double x0, x1, x2, xd;
                // column = corner c; rows in the reference's triplet order (solver.py:472-497)
                k0 = r4.y;
                k1 = r4.z;
                k2 = r4.w;
                if (c == 0) {
                    x0 = a12; x1 = a13; x2 = a14; xd = a11;
                } else if (c == 1) {
                    x0 = a12; x1 = a23; x2 = a24; xd = a22;
                } else if (c == 2) {
                    x0 = a23; x1 = a13; x2 = a34; xd = a33;
                } else {
                    x0 = a14; x1 = a24; x2 = a34; xd = a44;
                }
                if (want_pat) {
                    accumulate(keys, av, bv, cnt, k0, x0, bij, want_a, want_b);
                    accumulate(keys, av, bv, cnt, k1, x1, bij, want_a, want_b);
                    accumulate(keys, av, bv, cnt, k2, x2, bij, want_a, want_b);
                    accumulate(keys, av, bv, cnt, (int)r, xd, bii, want_a, want_b);
                }
            }
                };
        if (in_regs) {
            // fully unrolled: the element-record loads have register-known addresses, so the
            // compiler may hoist them above the shared-memory accumulation of earlier incidences
#pragma unroll
            for (int u = 0; u < RS; u++)
                if (u < ninc) body(qi[u], ldg_d4(rec + (qi[u].x >> 2)));
        } else {
            for (int p = beg; p < end; p++) {
                const int4 r4 = inc4[p];
                D4 q = {0.0, 0.0, 0.0, 0.0};
                if (K == 3) q = ldg_d4(rec + (r4.x >> 2));
                body(r4, q);
            }
        }
        if (out.lump_ptr && end > beg) {
            const int lp = out.lump_ptr[r];
            out.lump_idx[lp] = (int)r;
            out.lump_val[lp] = lump;
        }
        if (!use_smem && want_a && want_b) {
            for (int q = 0; q < cnt; q++) out.b_idx[rbeg + q] = keys[q];
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int q = threadIdx.x; q < blk_nnz; q += kRowThreads) {
            const int key = s_k[q];
            if (want_a) {
                out.a_idx[blk_beg + q] = key;
                out.a_val[blk_beg + q] = s_a[q];
            }
            if (want_b) {
                out.b_idx[blk_beg + q] = key;
                out.b_val[blk_beg + q] = s_b[q];
            }
        }
    }
}

// ---- fast triangle rows ---------------------------------------------------------------------------
// ncu (round 2) on the per-thread kernels above at level 9: count 172 us, fill 590 us, both bound by
// instruction issue - the sorted key set of a row was built by insertion into shared memory, one
// dependent compare / move at a time.  Fast form for rows with <= 8 incident triangles (every row of
// a regular surface mesh): the up to 16 neighbour ids of a row live in REGISTERS, packed with their
// position in the reference's triplet order into one 32-bit word (id << 4 | position), and are
// sorted by a 63-comparator odd-even merge network of min / max instructions - no branches, no
// memory.  A walk over the sorted words yields the unique keys (the row's column indices), the
// slot of the diagonal and, per triplet, the slot it adds into; the values are then accumulated in
// triplet order (caller's element id ascending), so every entry is bit-identical to the sequential
// model.  Rows with more incidences, a repeated vertex inside an element or ids >= 2^27 are flagged
// and done by the per-thread kernels restricted to them (only_rows).
#define LB_NET16(X)                                                                                              \
    X(0, 1) X(2, 3) X(0, 2) X(1, 3) X(1, 2) X(4, 5) X(6, 7) X(4, 6) X(5, 7) X(5, 6) X(0, 4) X(2, 6) X(2, 4)    \
    X(1, 5) X(3, 7) X(3, 5) X(1, 2) X(3, 4) X(5, 6) X(8, 9) X(10, 11) X(8, 10) X(9, 11) X(9, 10) X(12, 13)     \
    X(14, 15) X(12, 14) X(13, 15) X(13, 14) X(8, 12) X(10, 14) X(10, 12) X(9, 13) X(11, 15) X(11, 13)           \
    X(9, 10) X(11, 12) X(13, 14) X(0, 8) X(4, 12) X(4, 8) X(2, 10) X(6, 14) X(6, 10) X(2, 4) X(6, 8)            \
    X(10, 12) X(1, 9) X(5, 13) X(5, 9) X(3, 11) X(7, 15) X(7, 11) X(3, 5) X(7, 9) X(11, 13) X(1, 2) X(3, 4)     \
    X(5, 6) X(7, 8) X(9, 10) X(11, 12) X(13, 14)
#define LB_MINMAX(i, j)                      \
    {                                        \
        const int lo_ = min(w[i], w[j]);     \
        w[j] = max(w[i], w[j]);              \
        w[i] = lo_;                          \
    }
constexpr int kFastInc = 8;  // incidences per row handled in registers

__global__ void __launch_bounds__(kRowThreads) tria_row_count_fast(const int32_t *__restrict__ inc_ptr,
                                                                   const int4 *__restrict__ inc4, int64_t n,
                                                                   int32_t *__restrict__ row_nnz,
                                                                   int32_t *__restrict__ row_has,
                                                                   int32_t *__restrict__ row_big,
                                                                   int32_t *__restrict__ nbig) {
    const int64_t r = (int64_t)blockIdx.x * kRowThreads + threadIdx.x;
    if (r >= n) return;
    const int beg = inc_ptr[r], ninc = inc_ptr[r + 1] - beg;
    bool big = ninc > kFastInc || n >= (1ll << 27);
    int cnt = 0;
    if (!big && ninc > 0) {
        int w[16];
#pragma unroll
        for (int u = 0; u < kFastInc; u++) {
            int4 q = make_int4(0, INT_MAX, INT_MAX, 0);
            if (u < ninc) q = __ldg(inc4 + beg + u);
            w[2 * u] = q.y;
            w[2 * u + 1] = q.z;
        }
        LB_NET16(LB_MINMAX)
        int prev = -1;
        bool self = false;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int key = w[i];
            if (key != INT_MAX && key != prev) cnt++;
            self |= key == (int)r;
            prev = key;
        }
        cnt++;  // the diagonal
        big = self;  // a vertex repeated inside an element adds off-diagonal triplets to the diagonal
    }
    row_has[r] = ninc > 0;
    row_big[r] = big;
    row_nnz[r] = big ? 0 : cnt;  // flagged rows: row_count_kernel(only_rows) fills it in
    if (big) atomicAdd(nbig, 1);
}

__global__ void __launch_bounds__(kRowThreads) tria_row_fill_fast(
    const D4 *__restrict__ rec, const int32_t *__restrict__ inc_ptr, const int4 *__restrict__ inc4, int64_t n, int cap,
    const ElemConsts *__restrict__ consts, int degen_div_f32, RowOut out, const int32_t *__restrict__ row_big) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_a = reinterpret_cast<double *>(smem_raw);
    double *s_b = s_a + cap;
    int32_t *s_k = reinterpret_cast<int32_t *>(s_b + cap);
    unsigned char *s_slot = reinterpret_cast<unsigned char *>(s_k + cap);  // [16][T]: slot of triplet position c
    constexpr int T = kRowThreads;
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * T, r = r0 + t;
    const int64_t rlast = min(r0 + (int64_t)T, n);
    const int blk_beg = out.indptr[r0], blk_nnz = out.indptr[rlast] - blk_beg;
    const bool want_a = out.a_val != nullptr, want_b = out.b_val != nullptr;
    const bool want_pat = want_a || want_b;
    const bool use_smem = want_pat && blk_nnz <= cap;
    // the row image lives in shared memory (streamed out below) or, for an oversized block, in the
    // output itself: two instantiations of the same code so that the common case compiles to LDS / STS
    // instead of generic loads and stores
    auto process = [&](int32_t *keys, double *av, double *bv, const int rbeg, const int beg, const int ninc) {
        // 1. incidence records in the reference's triplet order (caller's element id, then corner)
        int4 qi[kFastInc];
#pragma unroll
        for (int u = 0; u < kFastInc; u++) qi[u] = u < ninc ? __ldg(inc4 + beg + u) : make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
#define LB_CSWAP(i, j)                                                              \
    if (qi[i].w > qi[j].w || (qi[i].w == qi[j].w && qi[i].x > qi[j].x)) {           \
        const int4 tmp_ = qi[i];                                                    \
        qi[i] = qi[j];                                                              \
        qi[j] = tmp_;                                                               \
    }
        LB_CSWAP(0, 1) LB_CSWAP(2, 3) LB_CSWAP(4, 5) LB_CSWAP(6, 7) LB_CSWAP(0, 2) LB_CSWAP(1, 3) LB_CSWAP(4, 6)
        LB_CSWAP(5, 7) LB_CSWAP(1, 2) LB_CSWAP(5, 6) LB_CSWAP(0, 4) LB_CSWAP(3, 7) LB_CSWAP(1, 5) LB_CSWAP(2, 6)
        LB_CSWAP(1, 4) LB_CSWAP(3, 6) LB_CSWAP(2, 4) LB_CSWAP(3, 5) LB_CSWAP(3, 4)
#undef LB_CSWAP
        // 2. element records: 8 independent 32-byte gathers (neighbouring rows share them: L1 / L2)
        D4 er[kFastInc];
#pragma unroll
        for (int u = 0; u < kFastInc; u++)
            if (u < ninc) er[u] = ldg_d4(rec + (qi[u].x >> 2));
        // 3. sort (neighbour id << 4 | triplet position); position c = 2u + s is the reference order
        int cnt = 0, dslot = 0;
        if (want_pat) {
            int w[16];
#pragma unroll
            for (int u = 0; u < kFastInc; u++) {
                w[2 * u] = u < ninc ? (qi[u].y << 4) | (2 * u) : INT_MAX;
                w[2 * u + 1] = u < ninc ? (qi[u].z << 4) | (2 * u + 1) : INT_MAX;
            }
            LB_NET16(LB_MINMAX)
            // 4. unique keys -> slots (bit 7: first triplet of its key); the diagonal takes the slot
            //    before the first larger key
            int prev = -1;
            bool dd = false;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (w[i] != INT_MAX) {
                    const int key = w[i] >> 4, cpos = w[i] & 15;
                    const bool first = key != prev;
                    if (first) {
                        if (!dd && key > (int)r) {
                            dslot = cnt++;
                            dd = true;
                        }
                        keys[cnt++] = key;
                        prev = key;
                    }
                    s_slot[cpos * T + t] = (unsigned char)((cnt - 1) | (first ? 0x80 : 0));
                }
            }
            if (!dd) dslot = cnt++;
            keys[dslot] = (int)r;
        }
        // 5. accumulate in triplet order; the first addend of a slot is stored as it is (keeps -0.0)
        double da = 0.0, db = 0.0, lump = 0.0;
#pragma unroll
        for (int u = 0; u < kFastInc; u++) {
            if (u < ninc) {
                const int c = qi[u].x & 3;
                double a12 = er[u].x, a23 = er[u].y, a31 = er[u].z, bii = er[u].w;
                if (bii < 0.0) {  // clamped element (solver.py:159): divide by the global mean
                    const double vm = consts->vol_mean;
                    if (degen_div_f32) {
                        a12 = (double)__fdiv_rn((float)a12, (float)vm);
                        a23 = (double)__fdiv_rn((float)a23, (float)vm);
                        a31 = (double)__fdiv_rn((float)a31, (float)vm);
                    } else {
                        a12 = __ddiv_rn(a12, vm);
                        a23 = __ddiv_rn(a23, vm);
                        a31 = __ddiv_rn(a31, vm);
                    }
                    bii = consts->bii_deg;
                }
                const double bij = 0.5 * bii;
                lump += 2.0 * bii;  // vol/12 (vol/3 for fem_tria_mass) == 2*bii exactly
                // column = corner c; rows in the reference's triplet order (solver.py:171-175)
                const double x0 = c == 2 ? a23 : a12, x1 = c == 1 ? a23 : a31;
                const double m0 = c == 2 ? a31 : a12, m1 = c == 0 ? a31 : a23;
                // the diagonal (row sum = 0, solver.py:167-169) is formed in the element dtype
                const double xd = degen_div_f32 ? (double)__fsub_rn(-(float)m0, (float)m1) : __dsub_rn(-m0, m1);
                if (want_pat) {
                    const int e0 = s_slot[(2 * u) * T + t], e1 = s_slot[(2 * u + 1) * T + t];
                    const int p0 = e0 & 0x7f, p1 = e1 & 0x7f;
                    if (want_a) av[p0] = (e0 & 0x80) ? x0 : __dadd_rn(av[p0], x0);
                    if (want_b) bv[p0] = (e0 & 0x80) ? bij : __dadd_rn(bv[p0], bij);
                    if (want_a) av[p1] = (e1 & 0x80) ? x1 : __dadd_rn(av[p1], x1);
                    if (want_b) bv[p1] = (e1 & 0x80) ? bij : __dadd_rn(bv[p1], bij);
                }
                da = u == 0 ? xd : __dadd_rn(da, xd);
                db = u == 0 ? bii : __dadd_rn(db, bii);
            }
        }
        if (want_pat) {
            if (want_a) av[dslot] = da;
            if (want_b) bv[dslot] = db;
            if (!use_smem && want_a && want_b)
                for (int q = 0; q < cnt; q++) out.b_idx[rbeg + q] = keys[q];
        }
        if (out.lump_ptr) {
            const int lp = out.lump_ptr[r];
            out.lump_idx[lp] = (int)r;
            out.lump_val[lp] = lump;
        }
    };
    if (r < n && !row_big[r]) {
        const int beg = inc_ptr[r], ninc = inc_ptr[r + 1] - beg;
        if (ninc > 0) {
            const int rbeg = out.indptr[r];
            if (use_smem || !want_pat) {
                const int off = want_pat ? rbeg - blk_beg : 0;
                process(s_k + off, s_a + off, s_b + off, rbeg, beg, ninc);
            } else {
                process((want_a ? out.a_idx : out.b_idx) + rbeg, out.a_val + rbeg, out.b_val + rbeg, rbeg, beg, ninc);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        // segments of flagged rows hold stale shared memory here; row_fill_kernel(only_rows) runs
        // after this kernel and rewrites them
        for (int q = t; q < blk_nnz; q += T) {
            const int key = s_k[q];
            if (want_a) {
                out.a_idx[blk_beg + q] = key;
                out.a_val[blk_beg + q] = s_a[q];
            }
            if (want_b) {
                out.b_idx[blk_beg + q] = key;
                out.b_val[blk_beg + q] = s_b[q];
            }
        }
    }
}

// ---- strip-cooperative triangle rows -----------------------------------------------------------
// ncu (round 2) on the pipeline above at level 9: 1.8 GB of DRAM traffic for 0.59 GB of compulsory
// bytes - 1.1 GB of it are intermediates the design itself creates (16-byte incidence records written
// once and read twice, 32-byte element records written and gathered) - and ~1600 instructions per
// row in the fill.  Strip form: a CTA owns 128 consecutive rows of the locality numbering.  The
// elements that touch those rows are (a) the contiguous range of the sorted element array whose
// smallest vertex lies in the strip and (b) a short list of "halo" elements whose smallest vertex
// lies in an earlier strip (lb_mesh::hptr / hlist, part of the solver layout built at upload).  The
// CTA loads them once, computes the element matrices INTO SHARED MEMORY, builds the strip-local
// vertex -> element incidence with shared-memory atomics, and every thread then merges one row out of
// shared memory with the same register sorting networks as tria_row_fill_fast.  No incidence or
// element records ever reach DRAM; the count pass (FILL = false) is the same traversal without
// the vertex loads and the arithmetic.
// Anything the fast path does not cover raises a flag and the whole assembly falls back to the
// record pipeline above: a row with > 8 incident triangles or a vertex repeated inside a triangle,
// a strip with > kStripElems elements, a degenerate element (its clamp needs the global mean of vol).
constexpr int kStripRows = 128;
constexpr int kStripThreads = 256;  // phase 1 (element loads + arithmetic) uses all of them, phase 2 one thread per row
constexpr int kStripElems = 576;  // measured: 410 at most on a level-8 icosphere (mean 334 = 256 own + 78 halo)

struct StripLayout {
    const int4 *t4m;
    const int32_t *kptr;   // (n + 1) first sorted element whose smallest vertex is >= v
    const int32_t *hptr;   // (nstrips + 1)
    const int32_t *hlist;  // halo elements per strip
    int64_t n;
};

// Single pass (FILL): the row pointers are not an input.  Every thread first sorts its row's neighbour
// keys (registers) and counts the distinct ones, a block scan turns the counts into offsets inside the
// strip, and the strip's first output position comes from a DECOUPLED LOOK-BACK over the strips (each CTA
// publishes its entry count, then its inclusive prefix, in a 64-bit state word; CTAs take their strip
// from an atomic ticket so that every predecessor is already running).  The kernel writes indptr itself;
// the output arrays are allocated from the entry count the COUNT pass (FILL = false) measured once at
// mesh upload (a topological constant of the mesh, like the number of its edges), and the total found
// here is checked against it.  Round 2: count 106 us + two scans + a host round trip + fill 404 us ->
// one kernel.
struct StripFused {
    unsigned long long *state;  // [nstrips] 0 = nothing yet; bit 63: inclusive prefix, bit 62: aggregate only
    int32_t *ticket;            // [1] next strip
    int32_t *indptr_a, *indptr_b;  // (n + 1) outputs (either may be NULL)
    int32_t *lump_ptr;          // (n + 1) output for a lumped mass (or NULL)
    int32_t *totals;            // [2] total entries / total rows with entries, written by the last strip
};

template <bool FILL, class T, int MODE>
__global__ void __launch_bounds__(kStripThreads, 4) strip_rows_kernel(
    StripLayout L, const typename Ex<T>::V4 *__restrict__ v4m, const double *__restrict__ u1,
    const double *__restrict__ u2, const double *__restrict__ am, int cap, RowOut out, StripFused fz,
    int32_t *__restrict__ row_nnz, int32_t *__restrict__ row_has, int32_t *__restrict__ flags) {
    constexpr int R = kStripRows;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    // layout: s_rec [E] D4 | s_el [E] int4 | s_a [cap] | s_b [cap] | s_k [cap] | s_cnt [R] | s_inc [8][R] u16 | s_slot [16][R] u8
    D4 *s_rec = reinterpret_cast<D4 *>(smem_raw);
    int4 *s_el = reinterpret_cast<int4 *>(s_rec + (FILL ? kStripElems : 0));
    double *s_a = reinterpret_cast<double *>(s_el + kStripElems);
    double *s_b = s_a + (FILL ? cap : 0);
    int32_t *s_k = reinterpret_cast<int32_t *>(s_b + (FILL ? cap : 0));
    int32_t *s_cnt = s_k + (FILL ? cap : 0);
    unsigned short *s_inc = reinterpret_cast<unsigned short *>(s_cnt + R);
    unsigned char *s_slot = reinterpret_cast<unsigned char *>(s_inc + 8 * R);
    __shared__ int s_strip, s_wsum[2][kStripThreads / 32], s_base[2];
    const int t = threadIdx.x;
    int strip = blockIdx.x;
    if (t < R) s_cnt[t] = 0;
    if (FILL && t == 0) s_strip = atomicAdd(fz.ticket, 1);
    __syncthreads();
    if (FILL) strip = s_strip;
    const int64_t r0 = (int64_t)strip * R, r = r0 + t;
    const int nrows = (int)(min(L.n, r0 + R) - r0);
    const int eb = L.kptr[r0], nown = L.kptr[r0 + nrows] - eb;
    const int hb = L.hptr[strip], ne = nown + (L.hptr[strip + 1] - hb);
    const bool oversized = ne > kStripElems;  // uniform
    if (oversized) {
        if (t == 0) atomicOr(flags, 1);
        if (!FILL) return;  // (FILL: the strip still has to take part in the look-back chain, with zero entries)
    }
    // ---- phase 1: elements -> shared memory, strip-local incidence
    for (int le = t; le < (oversized ? 0 : ne); le += kStripThreads) {
        const int e = le < nown ? eb + le : __ldg(L.hlist + hb + (le - nown));
        const int4 ti = __ldg(L.t4m + e);
        s_el[le] = ti;
        if (FILL) {
            const Vec3<T> p1 = load_vertex<T>(v4m, ti.x), p2 = load_vertex<T>(v4m, ti.y), p3 = load_vertex<T>(v4m, ti.z);
            double q12, q23, q31, bii, vol_d;
            tria_local<T, MODE>(p1, p2, p3, ti.w, u1, u2, am, q12, q23, q31, bii, vol_d);
            st_d4(s_rec + le, q12, q23, q31, bii);
            if (bii < 0.0) atomicOr(flags + 1, 1);  // degenerate: needs the global clamp -> record pipeline
        }
        const int vs[3] = {ti.x, ti.y, ti.z};
#pragma unroll
        for (int cn = 0; cn < 3; cn++) {
            const int64_t lv = (int64_t)vs[cn] - r0;
            if (lv >= 0 && lv < nrows) {
                const int slot = atomicAdd(&s_cnt[lv], 1);
                if (slot < kFastInc) s_inc[slot * R + lv] = (unsigned short)(le * 4 + cn);
            }
        }
    }
    __syncthreads();
    // ---- phase 2a: every thread sorts the neighbour keys of its row in registers and counts the distinct ones
    const bool want_a = FILL && out.a_val != nullptr, want_b = FILL && out.b_val != nullptr;
    const bool want_pat = want_a || want_b;
    const int ninc = (t < nrows && !oversized) ? s_cnt[t] : 0;
    const bool fast = ninc > 0 && ninc <= kFastInc;
    int code[kFastInc], w[16];
    int cnt = 0, dslot = 0;
    bool self = false;
    if (fast) {
        // incidences in the reference's triplet order: (caller's element id, corner)
        int w8[kFastInc];
#pragma unroll
        for (int u = 0; u < kFastInc; u++) {
            w8[u] = INT_MAX;
            if (u < ninc) {
                const int cd = s_inc[u * R + t];
                w8[u] = (s_el[cd >> 2].w << 5) | ((cd & 3) << 3) | u;
            }
        }
#define LB_MM8(i, j)                      \
    {                                     \
        const int lo_ = min(w8[i], w8[j]); \
        w8[j] = max(w8[i], w8[j]);         \
        w8[i] = lo_;                      \
    }
        LB_MM8(0, 1) LB_MM8(2, 3) LB_MM8(4, 5) LB_MM8(6, 7) LB_MM8(0, 2) LB_MM8(1, 3) LB_MM8(4, 6) LB_MM8(5, 7) LB_MM8(1, 2)
        LB_MM8(5, 6) LB_MM8(0, 4) LB_MM8(3, 7) LB_MM8(1, 5) LB_MM8(2, 6) LB_MM8(1, 4) LB_MM8(3, 6) LB_MM8(2, 4) LB_MM8(3, 5)
        LB_MM8(3, 4)
#undef LB_MM8
#pragma unroll
        for (int i = 0; i < kFastInc; i++) {
            code[i] = 0;
            w[2 * i] = w[2 * i + 1] = INT_MAX;
            if (i < ninc) {
                code[i] = s_inc[(w8[i] & 7) * R + t];
                const int4 el = s_el[code[i] >> 2];
                const int cn = code[i] & 3;
                // neighbours in the triplet order of this column: corner 0: (t2, t3), 1: (t1, t3), 2: (t2, t1)
                const int k0 = cn == 1 ? el.x : el.y, k1 = cn == 2 ? el.x : el.z;
                self |= k0 == (int)r || k1 == (int)r;
                w[2 * i] = (k0 << 4) | (2 * i);
                w[2 * i + 1] = (k1 << 4) | (2 * i + 1);
            }
        }
        LB_NET16(LB_MINMAX)
        int prev = -1;
        bool dd = false;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (w[i] != INT_MAX) {
                const int key = w[i] >> 4, cpos = w[i] & 15;
                const bool first = key != prev;
                if (first) {
                    if (!dd && key > (int)r) {
                        dslot = cnt++;
                        dd = true;
                    }
                    cnt++;
                    prev = key;
                }
                if (FILL) s_slot[cpos * R + t] = (unsigned char)((cnt - 1) | (first ? 0x80 : 0));
            }
        }
        if (!dd) dslot = cnt++;
        if (self) atomicOr(flags, 1);
    }
    if (t < nrows && ninc > kFastInc) atomicOr(flags, 1);
    if (!FILL) {
        if (t < nrows) {
            row_has[r] = ninc > 0;
            row_nnz[r] = fast ? cnt : 0;
        }
        return;
    }
    // ---- block scan of the row counts (and of "row has entries" for the lumped mass), look-back over the strips
    const int my_cnt = fast ? cnt : 0, my_has = ninc > 0 ? 1 : 0;
    int inc_cnt = my_cnt, inc_has = my_has;
    const int lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc_cnt, d), h = __shfl_up_sync(0xffffffffu, inc_has, d);
        if (lane >= d) {
            inc_cnt += a;
            inc_has += h;
        }
    }
    if (lane == 31) {
        s_wsum[0][wid] = inc_cnt;
        s_wsum[1][wid] = inc_has;
    }
    __syncthreads();
    int woff_cnt = 0, woff_has = 0, blk_nnz = 0, blk_has = 0;
#pragma unroll
    for (int k = 0; k < kStripThreads / 32; k++) {
        if (k < wid) {
            woff_cnt += s_wsum[0][k];
            woff_has += s_wsum[1][k];
        }
        blk_nnz += s_wsum[0][k];
        blk_has += s_wsum[1][k];
    }
    const int off = woff_cnt + inc_cnt - my_cnt, off_has = woff_has + inc_has - my_has;
    // state word: [63] inclusive prefix published, [62] aggregate published, [61:31] rows with entries, [30:0] entries
    constexpr unsigned long long kP = 1ull << 63, kA = 1ull << 62, kMask = (1ull << 62) - 1;
    const unsigned long long agg = ((unsigned long long)blk_has << 31) | (unsigned long long)blk_nnz;
    volatile unsigned long long *st = fz.state;
    // the strip's entry count is published at once (successors can sum over it), its own look-back runs
    // AFTER the rows have been merged into shared memory: by then the predecessors have published too.
    // (Looking back on one of the four row-less warps WHILE the others merge was measured 4-9 % slower:
    // that warp starts before the predecessors have published and spins on L2.)
    if (t == 0) st[strip] = (strip == 0 ? kP : kA) | agg;
    const bool use_smem = want_pat && blk_nnz <= cap;
    auto look_back = [&]() {  // warp 0; result in s_base
        unsigned long long excl = 0;
        if (strip > 0) {
            int j = strip - 1;  // predecessor window: lanes look at strips j, j-1, ..., j-31
            while (true) {
                const int idx = j - lane;
                unsigned long long v = 0;
                do {
                    v = idx >= 0 ? st[idx] : kP;  // strips before 0 count as a published prefix of zero
                } while (__any_sync(0xffffffffu, v == 0));
                const unsigned pmask = __ballot_sync(0xffffffffu, (v & kP) != 0);
                // sum the window up to and including the first published prefix (lowest lane with P)
                const int stop = pmask ? __ffs(pmask) - 1 : 31;
                unsigned long long part = (lane <= stop && idx >= 0) ? (v & kMask) : 0;
#pragma unroll
                for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                excl += part;
                if (pmask) break;
                j -= 32;
            }
            if (lane == 0) st[strip] = kP | (excl + agg);
        }
        if (lane == 0) {
            s_base[0] = (int)(excl & 0x7fffffffull);
            s_base[1] = (int)(excl >> 31);
            if (r0 + nrows == L.n) {  // the last strip closes the row pointers
                const int tot = s_base[0] + blk_nnz, toth = s_base[1] + blk_has;
                if (fz.indptr_a) fz.indptr_a[L.n] = tot;
                if (fz.indptr_b) fz.indptr_b[L.n] = tot;
                if (fz.lump_ptr) fz.lump_ptr[L.n] = toth;
                fz.totals[0] = tot;
                fz.totals[1] = toth;
            }
        }
    };
    if (!use_smem) {  // rows go straight to global memory (or only the lumped mass is built): positions needed now
        if (wid == 0) look_back();
        __syncthreads();
    }
    // ---- phase 2b: keys and values of the row into the shared-memory image of the strip's CSR segment
    int rbeg_direct = 0;
    // volatile: keeps the compiler from hoisting the load out of the branch (in the shared-memory case warp 0
    // writes s_base later, after its look-back; racecheck flags the speculative read although its value is unused)
    if (!use_smem) rbeg_direct = reinterpret_cast<volatile int *>(s_base)[0] + off;
    double lump = 0.0;
    if (fast) {
        int32_t *keys = use_smem ? s_k + off : (want_a ? out.a_idx : out.b_idx) + rbeg_direct;
        double *av = use_smem ? s_a + off : out.a_val + rbeg_direct;
        double *bv = use_smem ? s_b + off : out.b_val + rbeg_direct;
        if (want_pat) {
            int c2 = 0, prev = -1;
            bool dd = false;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (w[i] != INT_MAX) {
                    const int key = w[i] >> 4;
                    if (key != prev) {
                        if (!dd && key > (int)r) {
                            c2++;
                            dd = true;
                        }
                        keys[c2++] = key;
                        prev = key;
                    }
                }
            }
            keys[dslot] = (int)r;
        }
        double da = 0.0, db = 0.0;
#pragma unroll
        for (int i = 0; i < kFastInc; i++) {
            if (i < ninc) {
                const int cn = code[i] & 3;
                const D4 er = s_rec[code[i] >> 2];
                const double a12 = er.x, a23 = er.y, a31 = er.z, bii = er.w;
                const double bij = 0.5 * bii;
                lump += 2.0 * bii;  // vol/12 (vol/3 for fem_tria_mass) == 2*bii exactly
                const double x0 = cn == 2 ? a23 : a12, x1 = cn == 1 ? a23 : a31;
                const double m0 = cn == 2 ? a31 : a12, m1 = cn == 0 ? a31 : a23;
                // the diagonal (row sum = 0, solver.py:167-169) is formed in the element dtype
                const double xd = sizeof(T) == 4 && MODE != MODE_ANISO ? (double)__fsub_rn(-(float)m0, (float)m1) : __dsub_rn(-m0, m1);
                if (want_pat) {
                    const int e0 = s_slot[(2 * i) * R + t], e1 = s_slot[(2 * i + 1) * R + t];
                    const int p0 = e0 & 0x7f, p1 = e1 & 0x7f;
                    if (want_a) av[p0] = (e0 & 0x80) ? x0 : __dadd_rn(av[p0], x0);
                    if (want_b) bv[p0] = (e0 & 0x80) ? bij : __dadd_rn(bv[p0], bij);
                    if (want_a) av[p1] = (e1 & 0x80) ? x1 : __dadd_rn(av[p1], x1);
                    if (want_b) bv[p1] = (e1 & 0x80) ? bij : __dadd_rn(bv[p1], bij);
                }
                da = i == 0 ? xd : __dadd_rn(da, xd);
                db = i == 0 ? bii : __dadd_rn(db, bii);
            }
        }
        if (want_pat) {
            if (want_a) av[dslot] = da;
            if (want_b) bv[dslot] = db;
            if (!use_smem && want_a && want_b)
                for (int q = 0; q < cnt; q++) out.b_idx[rbeg_direct + q] = keys[q];
        }
    }
    if (use_smem) {  // warp 0 looks back once its own rows are merged; one barrier covers both
        if (wid == 0) look_back();
        __syncthreads();
    }
    const int blk_beg = s_base[0];
    if (t < nrows) {
        const int rbeg = blk_beg + off;
        if (fz.indptr_a) fz.indptr_a[r] = rbeg;
        if (fz.indptr_b) fz.indptr_b[r] = rbeg;
        if (fz.lump_ptr) {
            const int lp = s_base[1] + off_has;
            fz.lump_ptr[r] = lp;
            if (fast) {
                out.lump_idx[lp] = (int)r;
                out.lump_val[lp] = lump;
            }
        }
    }
    if (use_smem) {
        for (int q = t; q < blk_nnz; q += kStripThreads) {
            const int key = s_k[q];
            if (want_a) {
                out.a_idx[blk_beg + q] = key;
                out.a_val[blk_beg + q] = s_a[q];
            }
            if (want_b) {
                out.b_idx[blk_beg + q] = key;
                out.b_val[blk_beg + q] = s_b[q];
            }
        }
    }
}

#undef LB_MINMAX

// ---- fused tet rows (the default for tets: 9.4 -> 4.4 ms at 121^3 in the caller's numbering) --------
// ncu (round 1) on the per-thread tet kernels: count 1.5 ms + fill 6.9 ms at cube121, 9 % issue
// utilisation, 12 warps per SM, long-scoreboard bound - the fill sorts the row's ~24 incidence
// records in place in GLOBAL memory and then walks record -> element -> accumulate one dependent
// load at a time; the key set is discovered twice (count, fill).  Fused form, still one thread per
// row (SIMD across rows; a lane-cooperative variant was instruction bound):
//   A  row_accumulate_tet: sort (code, index) pairs in shared memory, gather 4 incidences = 4 + 12
//      independent loads at a time, accumulate (key, A, B) ONCE into a fixed-capacity slice of
//      shared memory ([slot][thread]: conflict free), write the block's image to a scratch of the
//      same layout fully coalesced, row_nnz = number of keys;
//   -  scan -> indptr (as before);
//   B  row_compact: scratch -> shared-memory image of the block's CSR segment -> coalesced stores.
// Rows with more than kFusedSI incidences or kFusedCap entries are flagged and done by the
// per-thread kernels restricted to those rows (only_rows).
constexpr int kFusedSI = 32;   // sortable incidences per row
constexpr int kFusedCap = 16;  // entries per row
constexpr int kFusedBatch = 4;

struct TetCorner {
    double x0, x1, x2, xd, bii, lmp;
};
// the four triplet values of corner c of a tet (column = corner, rows in the reference's triplet
// order solver.py:472-497), its mass entries and lumped share
__device__ __forceinline__ TetCorner tet_corner_values(int code, const D4 *__restrict__ rec,
                                                       const ElemConsts *__restrict__ consts, int degen_div_f32) {
    const int e = code >> 2, c = code & 3;
    const D4 *rp = rec + 3 * (int64_t)e;
    const D4 q0 = ldg_d4(rp), q1 = ldg_d4(rp + 1), q2 = ldg_d4(rp + 2);
    double a12 = q0.x, a13 = q0.y, a14 = q0.z, a23 = q0.w, a24 = q1.x, a34 = q1.y;
    double a11 = q1.z, a22 = q1.w, a33 = q2.x, a44 = q2.y, bii = q2.z, lmp = q2.w;
    if (bii < 0.0) {  // clamped element: divide the numerators by the global mean (solver.py:436-437)
        const double vm = consts->vol_mean;
        if (degen_div_f32) {
            const float vf = (float)vm;
            float f12 = __fdiv_rn((float)a12, vf), f13 = __fdiv_rn((float)a13, vf);
            float f14 = __fdiv_rn((float)a14, vf), f23 = __fdiv_rn((float)a23, vf);
            float f24 = __fdiv_rn((float)a24, vf), f34 = __fdiv_rn((float)a34, vf);
            a11 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f12, f13), f14), 6.f);
            a22 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f12, f23), f24), 6.f);
            a33 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f13, f23), f34), 6.f);
            a44 = (double)__fdiv_rn(__fsub_rn(__fsub_rn(-f14, f24), f34), 6.f);
            a12 = (double)__fdiv_rn(f12, 6.f); a13 = (double)__fdiv_rn(f13, 6.f);
            a14 = (double)__fdiv_rn(f14, 6.f); a23 = (double)__fdiv_rn(f23, 6.f);
            a24 = (double)__fdiv_rn(f24, 6.f); a34 = (double)__fdiv_rn(f34, 6.f);
        } else {
            double f12 = __ddiv_rn(a12, vm), f13 = __ddiv_rn(a13, vm), f14 = __ddiv_rn(a14, vm);
            double f23 = __ddiv_rn(a23, vm), f24 = __ddiv_rn(a24, vm), f34 = __ddiv_rn(a34, vm);
            a11 = __ddiv_rn(__dsub_rn(__dsub_rn(-f12, f13), f14), 6.0);
            a22 = __ddiv_rn(__dsub_rn(__dsub_rn(-f12, f23), f24), 6.0);
            a33 = __ddiv_rn(__dsub_rn(__dsub_rn(-f13, f23), f34), 6.0);
            a44 = __ddiv_rn(__dsub_rn(__dsub_rn(-f14, f24), f34), 6.0);
            a12 = __ddiv_rn(f12, 6.0); a13 = __ddiv_rn(f13, 6.0); a14 = __ddiv_rn(f14, 6.0);
            a23 = __ddiv_rn(f23, 6.0); a24 = __ddiv_rn(f24, 6.0); a34 = __ddiv_rn(f34, 6.0);
        }
        bii = consts->bii_deg;
        lmp = consts->lump_deg;
    }
    TetCorner t;
    t.bii = bii;
    t.lmp = lmp;
    if (c == 0) {
        t.x0 = a12; t.x1 = a13; t.x2 = a14; t.xd = a11;
    } else if (c == 1) {
        t.x0 = a12; t.x1 = a23; t.x2 = a24; t.xd = a22;
    } else if (c == 2) {
        t.x0 = a23; t.x1 = a13; t.x2 = a34; t.xd = a33;
    } else {
        t.x0 = a14; t.x1 = a24; t.x2 = a34; t.xd = a44;
    }
    return t;
}

constexpr size_t kFusedSmemA = (size_t)kRowThreads * (kFusedSI * 8 + kFusedCap * 20);

__global__ void __launch_bounds__(kRowThreads) row_accumulate_tet(
    const D4 *__restrict__ rec, const int32_t *__restrict__ inc_ptr, const int4 *__restrict__ inc4,
    const int32_t *__restrict__ eorig, int64_t n,
    const ElemConsts *__restrict__ consts, int degen_div_f32, int32_t *__restrict__ row_nnz,
    int32_t *__restrict__ row_has, int32_t *__restrict__ row_big, int32_t *__restrict__ sc_key,
    double *__restrict__ sc_a, double *__restrict__ sc_b, double *__restrict__ sc_lump) {
    constexpr int T = kRowThreads, SI = kFusedSI, CAP = kFusedCap, NB = kFusedBatch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *s_ord = reinterpret_cast<unsigned long long *>(smem_raw);  // [SI][T]
    double *s_a = reinterpret_cast<double *>(s_ord + SI * T);                      // [CAP][T]
    double *s_b = s_a + CAP * T;                                                   // [CAP][T]
    int32_t *s_k = reinterpret_cast<int32_t *>(s_b + CAP * T);                     // [CAP][T]
    __shared__ int s_maxcnt;
    const int t = threadIdx.x;
    const int64_t r = (int64_t)blockIdx.x * T + t;
    if (t == 0) s_maxcnt = 0;
    __syncthreads();
    int cnt = 0;
    bool big = false;
    if (r < n) {
        const int beg = inc_ptr[r], ninc = inc_ptr[r + 1] - beg;
        double lump = 0.0;
        if (ninc > SI) {
            big = true;
        } else if (ninc > 0) {
            // 1. (code << 8 | index) pairs, insertion-sorted in shared memory; codes fetched 8 at a time
            for (int i0 = 0; i0 < ninc; i0 += 8) {
                int c8[8], o8[8];
#pragma unroll
                for (int u = 0; u < 8; u++) c8[u] = i0 + u < ninc ? __ldg(&inc4[beg + i0 + u].x) : 0;
#pragma unroll
                for (int u = 0; u < 8; u++) o8[u] = i0 + u < ninc ? __ldg(eorig + (c8[u] >> 2)) : 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u;
                    if (i < ninc) {
                        // (caller's element id, corner) = the reference's triplet order; low 8 bits: record index
                        const unsigned long long key =
                            ((((unsigned long long)(unsigned)o8[u] << 2) | (unsigned)(c8[u] & 3)) << 8) | (unsigned)i;
                        int j = i - 1;
                        while (j >= 0 && s_ord[j * T + t] > key) {
                            s_ord[(j + 1) * T + t] = s_ord[j * T + t];
                            j--;
                        }
                        s_ord[(j + 1) * T + t] = key;
                    }
                }
            }
            // 2. replay in triplet order, NB incidences (NB + 3 NB independent loads) at a time
            auto acc = [&](int key, double va, double vb) {
                int pos = cnt;
                while (pos > 0 && s_k[(pos - 1) * T + t] >= key) pos--;
                if (pos < cnt && s_k[pos * T + t] == key) {
                    s_a[pos * T + t] = __dadd_rn(s_a[pos * T + t], va);
                    s_b[pos * T + t] = __dadd_rn(s_b[pos * T + t], vb);
                    return;
                }
                if (cnt == CAP) {
                    big = true;
                    return;
                }
                for (int q = cnt; q > pos; q--) {
                    s_k[q * T + t] = s_k[(q - 1) * T + t];
                    s_a[q * T + t] = s_a[(q - 1) * T + t];
                    s_b[q * T + t] = s_b[(q - 1) * T + t];
                }
                s_k[pos * T + t] = key;
                s_a[pos * T + t] = va;  // first addend itself, like csr_sum_duplicates (keeps -0.0)
                s_b[pos * T + t] = vb;
                cnt++;
            };
            for (int j0 = 0; j0 < ninc && !big; j0 += NB) {
                int4 q[NB];
                TetCorner tc[NB];
#pragma unroll
                for (int u = 0; u < NB; u++)
                    q[u] = j0 + u < ninc ? __ldg(inc4 + beg + (int)(s_ord[(j0 + u) * T + t] & 255ull)) : make_int4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < NB; u++)
                    if (j0 + u < ninc) tc[u] = tet_corner_values(q[u].x, rec, consts, degen_div_f32);
#pragma unroll
                for (int u = 0; u < NB; u++)
                    if (j0 + u < ninc) {
                        const double bij = 0.5 * tc[u].bii;
                        lump += tc[u].lmp;
                        acc(q[u].y, tc[u].x0, bij);
                        acc(q[u].z, tc[u].x1, bij);
                        acc(q[u].w, tc[u].x2, bij);
                        acc((int)r, tc[u].xd, tc[u].bii);
                    }
            }
        }
        row_has[r] = ninc > 0;
        row_big[r] = big;
        row_nnz[r] = big ? -1 : cnt;  // -1: the per-thread count kernel (only_rows) fills it in
        sc_lump[r] = lump;
        if (!big && cnt > 0) atomicMax(&s_maxcnt, cnt);
    }
    __syncthreads();
    // 3. the block's [slot][thread] image, slots below the block maximum, fully coalesced
    const int total = s_maxcnt * T;
    const size_t base = (size_t)blockIdx.x * CAP * T;
    for (int i = t; i < total; i += T) {
        sc_key[base + i] = s_k[i];
        sc_a[base + i] = s_a[i];
        sc_b[base + i] = s_b[i];
    }
}

// scratch ([block][slot][thread]) -> CSR.  Flagged rows are left to row_fill_kernel(only_rows).
__global__ void __launch_bounds__(kRowThreads) row_compact_kernel(int64_t n, const int32_t *__restrict__ row_big,
                                                                  const int32_t *__restrict__ sc_key,
                                                                  const double *__restrict__ sc_a,
                                                                  const double *__restrict__ sc_b,
                                                                  const double *__restrict__ sc_lump, RowOut out) {
    constexpr int T = kRowThreads, CAP = kFusedCap;
    __shared__ double s_a[CAP * T];
    __shared__ double s_b[CAP * T];
    __shared__ int32_t s_k[CAP * T];
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * T, r = r0 + t;
    const int64_t rlast = min(r0 + (int64_t)T, n);
    const int blk_beg = out.indptr[r0], blk_nnz = out.indptr[rlast] - blk_beg;
    const bool want_a = out.a_val != nullptr, want_b = out.b_val != nullptr;
    const bool use_smem = blk_nnz <= CAP * T;  // false only when flagged rows make the block long
    const size_t base = (size_t)blockIdx.x * CAP * T;
    if (r < n && !row_big[r]) {
        const int rbeg = out.indptr[r], cnt = out.indptr[r + 1] - rbeg;
        for (int q = 0; q < cnt; q++) {
            const int key = sc_key[base + q * T + t];
            const double a = sc_a[base + q * T + t], b = sc_b[base + q * T + t];
            if (use_smem) {
                const int o = rbeg - blk_beg + q;
                s_k[o] = key;
                s_a[o] = a;
                s_b[o] = b;
            } else {
                if (want_a) {
                    out.a_idx[rbeg + q] = key;
                    out.a_val[rbeg + q] = a;
                }
                if (want_b) {
                    out.b_idx[rbeg + q] = key;
                    out.b_val[rbeg + q] = b;
                }
            }
        }
        if (out.lump_ptr && out.lump_ptr[r + 1] > out.lump_ptr[r]) {
            const int lp = out.lump_ptr[r];
            out.lump_idx[lp] = (int)r;
            out.lump_val[lp] = sc_lump[r];
        }
    }
    if (use_smem) {
        __syncthreads();
        // segments of flagged rows hold stale shared memory here; row_fill_kernel(only_rows) runs
        // after this kernel and rewrites them
        for (int q = t; q < blk_nnz; q += T) {
            const int key = s_k[q];
            if (want_a) {
                out.a_idx[blk_beg + q] = key;
                out.a_val[blk_beg + q] = s_a[q];
            }
            if (want_b) {
                out.b_idx[blk_beg + q] = key;
                out.b_val[blk_beg + q] = s_b[q];
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------
static void build_incidence(lb_mesh *mesh, DBuf<int32_t> &deg) {
    // deg holds the per-vertex incidence histogram (from the element pass)
    lb_ctx *c = mesh->ctx;
    const int64_t nv = mesh->nv, nt = mesh->nt;
    mesh->inc_ptr.alloc(c, nv + 1);
    mesh->inc.alloc(c, (size_t)mesh->k * nt);
    exclusive_scan_i32(c, deg.p, mesh->inc_ptr.p, nv);
    deg.zero();
    LB_LAUNCH(c, incidence_fill, cdiv(nt, 256), 256, 0, mesh->t4.p, nt, mesh->k, mesh->inc_ptr.p, deg.p,
              mesh->inc.p);
    LB_LAUNCH(c, incidence_sort, cdiv(nv, 128), 128, 0, mesh->inc_ptr.p, mesh->inc.p, nv);
    mesh->has_inc = true;
}

__global__ void degree_count(const int4 *__restrict__ t4, int64_t nt, int k, int32_t *__restrict__ deg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    int4 ti = __ldg(t4 + e);
    atomicAdd(deg + ti.x, 1);
    atomicAdd(deg + ti.y, 1);
    atomicAdd(deg + ti.z, 1);
    if (k == 4) atomicAdd(deg + ti.w, 1);
}

// vertex -> element incidence without an assembly (used by the divergence gather)
void ensure_incidence(lb_mesh *mesh) {
    if (mesh->has_inc) return;
    lb_ctx *c = mesh->ctx;
    DBuf<int32_t> deg(c, mesh->nv);
    deg.zero();
    LB_LAUNCH(c, degree_count, cdiv(mesh->nt, 256), 256, 0, mesh->t4.p, mesh->nt, mesh->k, deg.p);
    build_incidence(mesh, deg);
}

// ---- locality ordering ----------------------------------------------------------------------------
// Vertices are binned into a 128^3 grid over the bounding box; bins are ordered along the Morton
// curve, vertices inside a bin by index.  Counting sort (histogram, scan, fill, per-bin sort), the
// same pattern as the incidence build; deterministic.
__global__ void bbox_kernel(const D4 *__restrict__ v4, int64_t n, double *__restrict__ out /* 6 x gridDim */) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        D4 p = ldg_d4(v4 + i);
        lo[0] = fmin(lo[0], p.x); hi[0] = fmax(hi[0], p.x);
        lo[1] = fmin(lo[1], p.y); hi[1] = fmax(hi[1], p.y);
        lo[2] = fmin(lo[2], p.z); hi[2] = fmax(hi[2], p.z);
    }
    __shared__ double red[6][256];
    for (int k = 0; k < 3; k++) {
        red[k][threadIdx.x] = lo[k];
        red[3 + k][threadIdx.x] = hi[k];
    }
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if (threadIdx.x < s)
            for (int k = 0; k < 3; k++) {
                red[k][threadIdx.x] = fmin(red[k][threadIdx.x], red[k][threadIdx.x + s]);
                red[3 + k][threadIdx.x] = fmax(red[3 + k][threadIdx.x], red[3 + k][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6) out[threadIdx.x * gridDim.x + blockIdx.x] = red[threadIdx.x][0];
}

__device__ __forceinline__ unsigned spread3(unsigned x) {  // 7 bits -> every third bit
    x &= 0x7f;
    x = (x | (x << 8)) & 0x0000700f;
    x = (x | (x << 4)) & 0x000430c3;
    x = (x | (x << 2)) & 0x00049249;
    return x;
}

__global__ void morton_cell_kernel(const D4 *__restrict__ v4, int64_t n, double ox, double oy, double oz, double sx,
                                   double sy, double sz, int32_t *__restrict__ cell, int32_t *__restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D4 p = ldg_d4(v4 + i);
    const unsigned ix = min(127, max(0, (int)((p.x - ox) * sx)));
    const unsigned iy = min(127, max(0, (int)((p.y - oy) * sy)));
    const unsigned iz = min(127, max(0, (int)((p.z - oz) * sz)));
    const int cid = (int)(spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2));
    cell[i] = cid;
    atomicAdd(hist + cid, 1);
}

__global__ void order_fill_kernel(int64_t n, const int32_t *__restrict__ cell, const int32_t *__restrict__ cptr,
                                  int32_t *__restrict__ cursor, int32_t *__restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cid = cell[i];
    order[cptr[cid] + atomicAdd(cursor + cid, 1)] = (int)i;
}

__global__ void invert_order_kernel(int64_t n, const int32_t *__restrict__ order, int32_t *__restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[order[i]] = (int)i;
}

void ensure_order(lb_order &o) {
    if (o.ready) return;
    lb_ctx *c = o.ctx;
    const int64_t n = o.n;
    const D4 *v4 = o.v4->p;
    constexpr int kCells = 128 * 128 * 128;
    const int nb = 256;
    DBuf<double> box(c, 6 * nb);
    LB_LAUNCH(c, bbox_kernel, nb, 256, 0, v4, n, box.p);
    std::vector<double> h(6 * nb);
    read_back(c, h.data(), box.p, h.size());
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = 1e300;
        hi[k] = -1e300;
        for (int b = 0; b < nb; b++) {
            lo[k] = std::min(lo[k], h[k * nb + b]);
            hi[k] = std::max(hi[k], h[(3 + k) * nb + b]);
        }
    }
    double sc[3];
    for (int k = 0; k < 3; k++) sc[k] = hi[k] > lo[k] ? 127.999 / (hi[k] - lo[k]) : 0.0;
    DBuf<int32_t> cell(c, n), hist(c, kCells), cptr(c, kCells + 1);
    hist.zero();
    LB_LAUNCH(c, morton_cell_kernel, cdiv(n, 256), 256, 0, v4, n, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], cell.p,
              hist.p);
    exclusive_scan_i32(c, hist.p, cptr.p, kCells);
    hist.zero();
    o.order.alloc(c, (size_t)n);
    o.inv.alloc(c, (size_t)n);
    LB_LAUNCH(c, order_fill_kernel, cdiv(n, 256), 256, 0, n, cell.p, cptr.p, hist.p, o.order.p);
    LB_LAUNCH(c, incidence_sort, cdiv(kCells, 128), 128, 0, cptr.p, o.order.p, (int64_t)kCells);
    LB_LAUNCH(c, invert_order_kernel, cdiv(n, 256), 256, 0, n, o.order.p, o.inv.p);
    o.ready = true;
}

// ---- solver layout of a mesh ---------------------------------------------------------------------
__global__ void gather_vertices_kernel(int64_t n, const int32_t *__restrict__ order, const D4 *__restrict__ v4,
                                       const float4 *__restrict__ v4f, D4 *__restrict__ v4m,
                                       float4 *__restrict__ v4fm) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int o = order[r];
    const D4 p = ldg_d4(v4 + o);
    st_d4(v4m + r, p.x, p.y, p.z, p.w);
    if (v4fm) v4fm[r] = __ldg(v4f + o);
}

// sort key of an element = its smallest new vertex id; histogram for the counting sort
__global__ void element_key_kernel(const int4 *__restrict__ t4, int64_t nt, int k, const int32_t *__restrict__ inv,
                                   int32_t *__restrict__ key, int32_t *__restrict__ hist) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    int m = min(inv[ti.x], min(inv[ti.y], inv[ti.z]));
    if (k == 4) m = min(m, inv[ti.w]);
    key[e] = m;
    atomicAdd(hist + m, 1);
}

__global__ void element_fill_kernel(int64_t nt, const int32_t *__restrict__ key, const int32_t *__restrict__ kptr,
                                    int32_t *__restrict__ cursor, int32_t *__restrict__ eperm) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int m = key[e];
    eperm[kptr[m] + atomicAdd(cursor + m, 1)] = (int)e;
}

// t4m[pos] = element eperm[pos] with new vertex ids; triangles carry the caller's element id in .w
__global__ void element_renumber_kernel(int64_t nt, int k, const int32_t *__restrict__ eperm,
                                        const int4 *__restrict__ t4, const int32_t *__restrict__ inv,
                                        int4 *__restrict__ t4m) {
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= nt) return;
    const int e = eperm[pos];
    const int4 ti = __ldg(t4 + e);
    t4m[pos] = make_int4(inv[ti.x], inv[ti.y], inv[ti.z], k == 4 ? inv[ti.w] : e);
}

// halo elements of the strips: element e (sorted, new vertex ids) belongs to the strip of its smallest
// vertex; every OTHER strip one of its vertices lies in gets it as a halo element.  ptr == NULL: count.
__global__ void halo_list_kernel(const int4 *__restrict__ t4m, int64_t nt, int32_t *__restrict__ cursor,
                                 const int32_t *__restrict__ ptr, int32_t *__restrict__ list) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4m + e);
    const int s0 = min(ti.x, min(ti.y, ti.z)) / kStripRows;
    const int sa = ti.x / kStripRows, sb = ti.y / kStripRows, sc = ti.z / kStripRows;
    const int cand[3] = {sa, sb, sc};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int sj = cand[i];
        bool dup = sj == s0;
        for (int j = 0; j < i; j++) dup |= cand[j] == sj;
        if (dup) continue;
        const int pos = atomicAdd(cursor + sj, 1);
        if (ptr) list[ptr[sj] + pos] = (int)e;
    }
}

static void refresh_layout_vertices(lb_mesh *m) {
    lb_ctx *c = m->ctx;
    const int64_t n = m->n_ref;
    const bool f32 = m->v_dtype == LB_F32;
    if (!m->v4m.p) m->v4m.alloc(c, n);
    if (f32 && !m->v4fm.p) m->v4fm.alloc(c, n);
    LB_LAUNCH(c, gather_vertices_kernel, cdiv(n, 256), 256, 0, n, m->ord->order.p, m->v4s->p,
              f32 ? m->v4f.p : (const float4 *)nullptr, m->v4m.p, f32 ? m->v4fm.p : (float4 *)nullptr);
}

// Locality numbering + renumbered, sorted element array.  Deterministic: elements with the same key
// are ordered by their id (per-key insertion sort behind the atomic fill).
static void build_layout(lb_mesh *m) {
    lb_ctx *c = m->ctx;
    const int64_t n = m->n_ref, nt = m->nt;
    m->ord = std::make_shared<lb_order>();
    m->ord->ctx = c;
    m->ord->v4 = m->v4s;
    m->ord->n = n;
    ensure_order(*m->ord);
    refresh_layout_vertices(m);
    DBuf<int32_t> key(c, nt), hist(c, n);
    m->kptr.alloc(c, n + 1);
    hist.zero();
    LB_LAUNCH(c, element_key_kernel, cdiv(nt, 256), 256, 0, m->t4.p, nt, m->k, m->ord->inv.p, key.p, hist.p);
    exclusive_scan_i32(c, hist.p, m->kptr.p, n);
    hist.zero();
    m->eorig.alloc(c, nt);
    LB_LAUNCH(c, element_fill_kernel, cdiv(nt, 256), 256, 0, nt, key.p, m->kptr.p, hist.p, m->eorig.p);
    LB_LAUNCH(c, incidence_sort, cdiv(n, 128), 128, 0, m->kptr.p, m->eorig.p, n);
    m->t4m.alloc(c, nt);
    LB_LAUNCH(c, element_renumber_kernel, cdiv(nt, 256), 256, 0, nt, m->k, m->eorig.p, m->t4.p, m->ord->inv.p, m->t4m.p);
    // triangles: per strip of 128 rows the "halo" elements (they touch the strip, their smallest vertex
    // lies in an earlier strip); ids fit the packed words of the strip kernel below 2^26 elements / 2^27 rows
    m->has_strips = false;
    if (m->k == 3 && nt < (1ll << 26) && n < (1ll << 27)) {
        const int ns = cdiv(n, kStripRows);
        DBuf<int32_t> hcnt(c, ns);
        hcnt.zero();
        m->hptr.alloc(c, ns + 1);
        LB_LAUNCH(c, halo_list_kernel, cdiv(nt, 256), 256, 0, m->t4m.p, nt, hcnt.p, (const int32_t *)nullptr, (int32_t *)nullptr);
        exclusive_scan_i32(c, hcnt.p, m->hptr.p, ns);
        int32_t total = 0;
        read_back(c, &total, m->hptr.p + ns, 1);
        m->hlist.alloc(c, (size_t)std::max(1, total));
        hcnt.zero();
        LB_LAUNCH(c, halo_list_kernel, cdiv(nt, 256), 256, 0, m->t4m.p, nt, hcnt.p, m->hptr.p, m->hlist.p);
        LB_LAUNCH(c, incidence_sort, cdiv(ns, 128), 128, 0, m->hptr.p, m->hlist.p, (int64_t)ns);  // deterministic order
        m->has_strips = true;
        // Topological constants of the mesh for the single-pass assembly: the number of stored entries of
        // its operators (vertices with an element + 2 x edges) and of rows with an element, plus whether
        // every row qualifies for the strip kernel.  The assembly sizes its outputs with them.
        {
            StripLayout L{m->t4m.p, m->kptr.p, m->hptr.p, m->hlist.p, n};
            DBuf<int32_t> row_nnz(c, n), row_has(c, n), flags(c, 2), scan(c, n + 1);
            flags.zero();
            const size_t smem_count = (size_t)kStripElems * 16 + kStripRows * 4 + 8 * kStripRows * 2;
            LB_LAUNCH(c, (strip_rows_kernel<false, double, MODE_FEM>), ns, kStripThreads, smem_count, L, (const D4 *)nullptr,
                      (const double *)nullptr, (const double *)nullptr, (const double *)nullptr, 0, RowOut{}, StripFused{},
                      row_nnz.p, row_has.p, flags.p);
            int32_t tot[2] = {0, 0}, hflags[2] = {0, 0};
            exclusive_scan_i32(c, row_nnz.p, scan.p, n);
            read_back(c, &tot[0], scan.p + n, 1);
            exclusive_scan_i32(c, row_has.p, scan.p, n);
            read_back(c, &tot[1], scan.p + n, 1);
            read_back(c, hflags, flags.p, 2);
            m->strip_fast = hflags[0] == 0;
            m->strip_nnz = tot[0];
            m->strip_nlump = tot[1];
        }
    }
}

template <class T>
static void run_element_pass(lb_mesh *mesh, int kind, const double *u1, const double *u2, const double *am,
                             D4 *rec, int32_t *deg, ElemConsts *consts) {
    lb_ctx *c = mesh->ctx;
    const int64_t nt = mesh->nt;
    const int nblocks = cdiv(nt, 256);
    DBuf<double> partial(c, nblocks);
    const typename Ex<T>::V4 *v4;
    if constexpr (sizeof(T) == 4) v4 = mesh->v4fm.p;
    else v4 = mesh->v4m.p;
    switch (kind) {
        case LB_FEM_TRIA: {
            auto kern = tria_element_kernel<T, MODE_FEM>;
            LB_LAUNCH(c, kern, nblocks, 256, 0, v4, mesh->t4m.p, nt, u1, u2, am, rec, deg, partial.p);
            break;
        }
        case LB_FEM_TRIA_ANISO: {
            auto kern = tria_element_kernel<T, MODE_ANISO>;
            LB_LAUNCH(c, kern, nblocks, 256, 0, v4, mesh->t4m.p, nt, u1, u2, am, rec, deg, partial.p);
            break;
        }
        case LB_FEM_TRIA_MASS: {
            auto kern = tria_element_kernel<T, MODE_MASS>;
            LB_LAUNCH(c, kern, nblocks, 256, 0, v4, mesh->t4m.p, nt, u1, u2, am, rec, deg, partial.p);
            break;
        }
        default: {
            auto kern = tet_element_kernel<T>;
            LB_LAUNCH(c, kern, nblocks, 256, 0, v4, mesh->t4m.p, nt, rec, deg, partial.p);
        }
    }
    auto fin = finalize_consts<T>;
    LB_LAUNCH(c, fin, 1, 256, 0, partial.p, nblocks, nt, kind, consts);
}

static lb_mat *new_mat(lb_ctx *c, int64_t n, int64_t nnz) {
    lb_mat *m = new lb_mat();
    m->ctx = c;
    m->n = n;
    m->nnz = nnz;
    m->indptr.alloc(c, n + 1);
    m->indices.alloc(c, nnz);
    m->data.alloc(c, nnz);
    return m;
}

template <int K>
static void run_rows(lb_mesh *mesh, const D4 *rec, const ElemConsts *consts, const int32_t *aptr, int4 *inc4,
                     bool want_a, bool lump, bool degen_f32, lb_mat **a_out, lb_mat **b_out) {
    lb_ctx *c = mesh->ctx;
    const int64_t n = mesh->n_ref;
    const int nblocks = cdiv(n, kRowThreads);
    const int32_t *eorig = mesh->eorig.p;
    // --- count
    const int cap_keys = K == 3 ? 4096 : 12288;  // int32 keys of shared memory per block
    DBuf<int32_t> row_nnz(c, n), row_has(c, n);
    DBuf<int32_t> scratch(c, (size_t)mesh->k * mesh->nt * (K - 1) + n);
    if (cap_keys * 4 > 40 * 1024)
        LB_CUDA(cudaFuncSetAttribute(row_count_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap_keys * 4));
    // tets: accumulate once into a per-row scratch, compact after the scan (row_accumulate_tet); rows it
    // flags (> 32 incidences or > 16 entries) go through the per-thread kernels restricted to them
    const bool fused = K == 4 && (want_a || !lump);
    DBuf<int32_t> row_big, sc_key;
    DBuf<double> sc_a, sc_b, sc_lump;
    if (fused) {
        const size_t slots = (size_t)nblocks * kFusedCap * kRowThreads;
        row_big.alloc(c, n);
        sc_key.alloc(c, slots);
        sc_a.alloc(c, slots);
        sc_b.alloc(c, slots);
        sc_lump.alloc(c, n);
        LB_CUDA(cudaFuncSetAttribute(row_accumulate_tet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemA));
        LB_LAUNCH(c, row_accumulate_tet, nblocks, kRowThreads, kFusedSmemA, rec, aptr, inc4, eorig, n, consts,
                  (int)degen_f32, row_nnz.p, row_has.p, row_big.p, sc_key.p, sc_a.p, sc_b.p, sc_lump.p);
    }
    // triangles: register fast path (tria_row_count_fast / tria_row_fill_fast), flagged rows by the
    // per-thread kernels
    const bool fast3 = K == 3;
    DBuf<int32_t> nbig(c, 1);
    if (fast3) {
        row_big.alloc(c, n);
        nbig.zero();
        LB_LAUNCH(c, tria_row_count_fast, nblocks, kRowThreads, 0, aptr, inc4, n, row_nnz.p, row_has.p, row_big.p, nbig.p);
    }
    LB_LAUNCH(c, row_count_kernel<K>, nblocks, kRowThreads, cap_keys * 4, aptr, inc4, n, cap_keys, scratch.p,
              row_nnz.p, row_has.p, (fused || fast3) ? row_big.p : (const int32_t *)nullptr);
    phase(c, "row count");
    const bool full_b = !lump;
    lb_mat *A = nullptr, *B = nullptr;
    DBuf<int32_t> indptr(c, n + 1), lump_ptr;
    exclusive_scan_i32(c, row_nnz.p, indptr.p, n);
    int32_t totals[2] = {0, 0};
    if (lump) {
        lump_ptr.alloc(c, n + 1);
        exclusive_scan_i32(c, row_has.p, lump_ptr.p, n);
        read_back(c, &totals[1], lump_ptr.p + n, 1);
    }
    read_back(c, &totals[0], indptr.p + n, 1);
    int32_t h_nbig = 0;
    if (fast3) read_back(c, &h_nbig, nbig.p, 1);
    const int64_t nnz = totals[0];
    phase(c, "scan + nnz readback");
    try {
        RowOut out{};
        out.indptr = indptr.p;
        if (want_a) {
            A = new_mat(c, n, nnz);
            d2d(c, A->indptr.p, indptr.p, (n + 1) * sizeof(int32_t));
            out.a_idx = A->indices.p;
            out.a_val = A->data.p;
        }
        if (full_b) {
            B = new_mat(c, n, nnz);
            d2d(c, B->indptr.p, indptr.p, (n + 1) * sizeof(int32_t));
            out.b_idx = B->indices.p;
            out.b_val = B->data.p;
        } else {
            B = new_mat(c, n, totals[1]);
            B->diagonal = true;
            d2d(c, B->indptr.p, lump_ptr.p, (n + 1) * sizeof(int32_t));
            out.lump_ptr = lump_ptr.p;
            out.lump_idx = B->indices.p;
            out.lump_val = B->data.p;
        }
        phase(c, "alloc outputs");
        if (want_a || full_b) {
            const int cap = K == 3 ? 1792 : 3072;  // CSR entries of shared memory per block
            const int smem = cap * 20;
            if (smem > 40 * 1024)
                LB_CUDA(cudaFuncSetAttribute(row_fill_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            if (fast3) {
                const int smem_fast = cap * 20 + 16 * kRowThreads;
                LB_LAUNCH(c, tria_row_fill_fast, nblocks, kRowThreads, smem_fast, rec, aptr, inc4, n, cap, consts,
                          (int)degen_f32, out, (const int32_t *)row_big.p);
                if (h_nbig > 0)
                    LB_LAUNCH(c, row_fill_kernel<K>, nblocks, kRowThreads, 16, rec, aptr, inc4, eorig, n, cap, consts,
                              (int)degen_f32, out, (const int32_t *)row_big.p);
            } else if (fused) {
                LB_LAUNCH(c, row_compact_kernel, nblocks, kRowThreads, 0, n, row_big.p, sc_key.p, sc_a.p, sc_b.p, sc_lump.p,
                          out);
                LB_LAUNCH(c, row_fill_kernel<K>, nblocks, kRowThreads, 16, rec, aptr, inc4, eorig, n, cap, consts,
                          (int)degen_f32, out, (const int32_t *)row_big.p);
            } else {
                LB_LAUNCH(c, row_fill_kernel<K>, nblocks, kRowThreads, smem, rec, aptr, inc4, eorig, n, cap, consts,
                          (int)degen_f32, out, (const int32_t *)nullptr);
            }
        } else {
            // mass only + lumped: no CSR pattern needed, but the same kernels do the sums
            const int cap = 0;
            if (fast3) {
                LB_LAUNCH(c, tria_row_fill_fast, nblocks, kRowThreads, 16 * kRowThreads, rec, aptr, inc4, n, cap, consts,
                          (int)degen_f32, out, (const int32_t *)row_big.p);
                if (h_nbig > 0)
                    LB_LAUNCH(c, row_fill_kernel<K>, nblocks, kRowThreads, 16, rec, aptr, inc4, eorig, n, cap, consts,
                              (int)degen_f32, out, (const int32_t *)row_big.p);
            } else {
                LB_LAUNCH(c, row_fill_kernel<K>, nblocks, kRowThreads, 16, rec, aptr, inc4, eorig, n, cap, consts,
                          (int)degen_f32, out, (const int32_t *)nullptr);
            }
        }
    } catch (...) {
        delete A;
        delete B;
        throw;
    }
    if (a_out) *a_out = A;
    else delete A;
    *b_out = B;
}

// strip-cooperative triangle assembly (strip_rows_kernel): count pass, scan, fill pass.  Returns false
// - with nothing allocated - when the mesh needs the record pipeline (flags raised by the kernels).
template <class T>
static bool run_strip_rows(lb_mesh *mesh, int kind, const double *u1, const double *u2, const double *am, bool want_a,
                           bool lump, lb_mat **a_out, lb_mat **b_out) {
    lb_ctx *c = mesh->ctx;
    const int64_t n = mesh->n_ref;
    const int ns = cdiv(n, kStripRows);
    StripLayout L{mesh->t4m.p, mesh->kptr.p, mesh->hptr.p, mesh->hlist.p, n};
    const typename Ex<T>::V4 *v4;
    if constexpr (sizeof(T) == 4) v4 = mesh->v4fm.p;
    else v4 = mesh->v4m.p;
    if (!mesh->strip_fast) return false;  // a row / strip the fast path does not cover (measured at upload)
    const int64_t nnz = mesh->strip_nnz, nlump = mesh->strip_nlump;
    DBuf<int32_t> flags(c, 2), ticket(c, 1), totals(c, 2);
    DBuf<unsigned long long> state(c, ns);
    flags.zero();
    ticket.zero();
    totals.zero();
    state.zero();
    const bool full_b = !lump;
    lb_mat *A = nullptr, *B = nullptr;
    try {
        RowOut out{};
        StripFused fz{};
        fz.state = state.p;
        fz.ticket = ticket.p;
        fz.totals = totals.p;
        if (want_a) {
            A = new_mat(c, n, nnz);
            fz.indptr_a = A->indptr.p;
            out.a_idx = A->indices.p;
            out.a_val = A->data.p;
        }
        if (full_b) {
            B = new_mat(c, n, nnz);
            fz.indptr_b = B->indptr.p;
            out.b_idx = B->indices.p;
            out.b_val = B->data.p;
        } else {
            B = new_mat(c, n, nlump);
            B->diagonal = true;
            fz.lump_ptr = B->indptr.p;
            out.lump_idx = B->indices.p;
            out.lump_val = B->data.p;
        }
        const int cap = (want_a || full_b) ? 1152 : 0;  // CSR entries of shared memory per strip (9 per row)
        const size_t smem = (size_t)kStripElems * 48 + (size_t)cap * 20 + kStripRows * 4 + 8 * kStripRows * 2 + 16 * kStripRows;
#define LB_STRIP(MODE)                                                                                                   \
    do {                                                                                                                 \
        LB_CUDA(cudaFuncSetAttribute(strip_rows_kernel<true, T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        LB_LAUNCH(c, (strip_rows_kernel<true, T, MODE>), ns, kStripThreads, smem, L, v4, u1, u2, am, cap, out, fz,       \
                  (int32_t *)nullptr, (int32_t *)nullptr, flags.p);                                                      \
    } while (0)
        if (kind == LB_FEM_TRIA) LB_STRIP(MODE_FEM);
        else if (kind == LB_FEM_TRIA_ANISO) LB_STRIP(MODE_ANISO);
        else LB_STRIP(MODE_MASS);
#undef LB_STRIP
        // one read-back at the end: flags (a degenerate element needs the global mean of vol -> record
        // pipeline) and the totals the look-back arrived at (must equal the mesh's constants)
        int32_t hflags[2] = {0, 0}, htot[2] = {0, 0};
        d2h(c, c->pinned, flags.p, 2 * sizeof(int32_t));
        d2h(c, (char *)c->pinned + 8, totals.p, 2 * sizeof(int32_t));
        sync(c);
        std::memcpy(hflags, c->pinned, 8);
        std::memcpy(htot, (char *)c->pinned + 8, 8);
        if (hflags[0] || hflags[1] || htot[0] != nnz || htot[1] != nlump) {
            delete A;
            delete B;
            return false;
        }
    } catch (...) {
        delete A;
        delete B;
        throw;
    }
    if (a_out) *a_out = A;
    else delete A;
    *b_out = B;
    return true;
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_mesh_create(lb_ctx *c, const void *v, int v_dtype, int64_t nv, const void *t, int t_itemsize,
                   int64_t nt, int k, lb_mesh **out) {
    LB_API_BEGIN
    LB_REQUIRE(c && v && t && out, "lb_mesh_create: NULL argument");
    LB_REQUIRE(k == 3 || k == 4, "elements must have 3 or 4 vertices, got %d", k);
    LB_REQUIRE(v_dtype == LB_F32 || v_dtype == LB_F64, "vertex dtype must be float32 or float64");
    LB_REQUIRE(t_itemsize == 4 || t_itemsize == 8, "element indices must be int32 or int64");
    LB_REQUIRE(nv > 0 && nt > 0, "empty mesh");
    LB_REQUIRE(nv < (1ll << 31) - 1 && nt < (1ll << 29), "mesh too large for int32 indexing");
    DeviceGuard g(c->device);
    lb_mesh *m = new lb_mesh();
    try {
        m->ctx = c;
        m->nv = nv;
        m->nt = nt;
        m->k = k;
        m->v_dtype = v_dtype;
        m->v4s = std::make_shared<DBuf<D4>>(c, (size_t)nv);
        m->t4.alloc(c, nt);
        if (v_dtype == LB_F32) m->v4f.alloc(c, nv);
        const size_t vbytes = (size_t)nv * 3 * (v_dtype == LB_F32 ? 4 : 8);
        const size_t tbytes = (size_t)nt * k * t_itemsize;
        DBuf<unsigned char> raw_v(c, vbytes), raw_t(c, tbytes);
        DBuf<long long> minmax(c, 2);
        long long init[2] = {LLONG_MAX, LLONG_MIN};
        std::memcpy(c->pinned, init, sizeof(init));
        h2d(c, minmax.p, c->pinned, sizeof(init));
        h2d(c, raw_v.p, v, vbytes);
        h2d(c, raw_t.p, t, tbytes);
        if (v_dtype == LB_F32)
            LB_LAUNCH(c, convert_vertices<float>, cdiv(nv, 256), 256, 0, (const float *)raw_v.p, nv, m->v4s->p, m->v4f.p);
        else
            LB_LAUNCH(c, convert_vertices<double>, cdiv(nv, 256), 256, 0, (const double *)raw_v.p, nv, m->v4s->p,
                      (float4 *)nullptr);
        if (t_itemsize == 4)
            LB_LAUNCH(c, convert_elements<int32_t>, cdiv(nt, 256), 256, 0, (const int32_t *)raw_t.p, nt, k, m->t4.p,
                      minmax.p);
        else
            LB_LAUNCH(c, convert_elements<int64_t>, cdiv(nt, 256), 256, 0, (const int64_t *)raw_t.p, nt, k, m->t4.p,
                      minmax.p);
        long long mm[2];
        read_back(c, mm, minmax.p, 2);
        LB_REQUIRE(mm[0] >= 0, "negative vertex index %lld in elements", mm[0]);
        LB_REQUIRE(mm[1] < nv, "Max index exceeds number of vertices");
        m->n_ref = mm[1] + 1;
        build_layout(m);
        sync(c);
    } catch (...) {
        delete m;
        throw;
    }
    *out = m;
    LB_API_END
}

int lb_mesh_update_vertices(lb_mesh *m, const void *v, int v_dtype) {
    LB_API_BEGIN
    LB_REQUIRE(m && v, "NULL argument");
    LB_REQUIRE(v_dtype == LB_F32 || v_dtype == LB_F64, "vertex dtype must be float32 or float64");
    lb_ctx *c = m->ctx;
    DeviceGuard g(c->device);
    const size_t vbytes = (size_t)m->nv * 3 * (v_dtype == LB_F32 ? 4 : 8);
    DBuf<unsigned char> raw_v(c, vbytes);
    h2d(c, raw_v.p, v, vbytes);
    m->v_dtype = v_dtype;
    if (v_dtype == LB_F32) {
        if (!m->v4f.p) m->v4f.alloc(c, m->nv);
        LB_LAUNCH(c, convert_vertices<float>, cdiv(m->nv, 256), 256, 0, (const float *)raw_v.p, m->nv, m->v4s->p,
                  m->v4f.p);
    } else {
        LB_LAUNCH(c, convert_vertices<double>, cdiv(m->nv, 256), 256, 0, (const double *)raw_v.p, m->nv, m->v4s->p,
                  (float4 *)nullptr);
    }
    refresh_layout_vertices(m);  // same numbering (a permutation stays valid when vertices move)
    sync(c);  // raw_v is borrowed from the caller until here
    LB_API_END
}

int lb_mesh_drop_cache(lb_mesh *m) {
    LB_API_BEGIN
    LB_REQUIRE(m, "mesh is NULL");
    DeviceGuard g(m->ctx->device);
    m->inc_ptr.release();
    m->inc.release();
    m->has_inc = false;  // the solver layout (numbering, sorted elements) is part of the upload and stays
    LB_API_END
}

int lb_mesh_free(lb_mesh *m) {
    LB_API_BEGIN
    if (!m) return LB_OK;
    DeviceGuard g(m->ctx->device);
    delete m;
    LB_API_END
}

int lb_fem_assemble(lb_ctx *c, lb_mesh *mesh, int kind, int lump, const double *u1, const double *u2,
                    const double *aniso_mat, lb_mat **a_out, lb_mat **b_out) {
    LB_API_BEGIN
    LB_REQUIRE(c && mesh && b_out, "lb_fem_assemble: NULL argument");
    LB_REQUIRE(mesh->ctx == c, "mesh belongs to another context");
    LB_REQUIRE(kind >= LB_FEM_TRIA && kind <= LB_FEM_TETRA, "unknown operator kind %d", kind);
    LB_REQUIRE((kind == LB_FEM_TETRA) == (mesh->k == 4), "operator kind does not match element type");
    if (kind == LB_FEM_TRIA_ANISO) LB_REQUIRE(u1 && u2 && aniso_mat, "anisotropic operator needs u1, u2, aniso_mat");
    DeviceGuard g(c->device);
    phase(c, "(enter assemble)");
    const int64_t nt = mesh->nt;
    const bool want_a = kind != LB_FEM_TRIA_MASS && a_out != nullptr;
    DBuf<double> d_u1, d_u2, d_am;
    if (kind == LB_FEM_TRIA_ANISO) {
        d_u1.alloc(c, 3 * nt);
        d_u2.alloc(c, 3 * nt);
        d_am.alloc(c, 2 * nt);
        h2d(c, d_u1.p, u1, 3 * nt * sizeof(double));
        h2d(c, d_u2.p, u2, 3 * nt * sizeof(double));
        h2d(c, d_am.p, aniso_mat, 2 * nt * sizeof(double));
    }
    if (a_out) *a_out = nullptr;
    auto tag = [&]() {  // the matrices are stored in the mesh's locality numbering (lb_mat::permuted)
        if (a_out && *a_out) {
            (*a_out)->ord = mesh->ord;
            (*a_out)->permuted = true;
        }
        (*b_out)->ord = mesh->ord;
        (*b_out)->permuted = true;
    };
    // triangles: strip-cooperative kernels (no record arrays in DRAM); anything they do not cover
    // (valence > 8, repeated vertex, degenerate element, oversized strip) falls through to the record pipeline
    if (mesh->k == 3 && mesh->has_strips) {
        const bool ok = mesh->v_dtype == LB_F32
                            ? run_strip_rows<float>(mesh, kind, d_u1.p, d_u2.p, d_am.p, want_a, lump != 0, a_out, b_out)
                            : run_strip_rows<double>(mesh, kind, d_u1.p, d_u2.p, d_am.p, want_a, lump != 0, a_out, b_out);
        phase(c, ok ? "strip rows" : "strip rows (fell back)");
        if (ok) {
            tag();
            c->n_strip_assemblies++;
            sync(c);  // u1/u2/aniso_mat are borrowed host buffers
            return LB_OK;
        }
    }
    c->n_record_assemblies++;
    DBuf<D4> rec(c, (size_t)nt * (mesh->k == 4 ? 3 : 1));
    DBuf<ElemConsts> consts(c, 1);
    DBuf<int32_t> deg(c, mesh->n_ref), aptr(c, mesh->n_ref + 1);
    DBuf<int4> inc4(c, (size_t)mesh->k * nt);
    deg.zero();
    if (mesh->v_dtype == LB_F32)
        run_element_pass<float>(mesh, kind, d_u1.p, d_u2.p, d_am.p, rec.p, deg.p, consts.p);
    else
        run_element_pass<double>(mesh, kind, d_u1.p, d_u2.p, d_am.p, rec.p, deg.p, consts.p);
    phase(c, "element pass");
    exclusive_scan_i32(c, deg.p, aptr.p, mesh->n_ref);
    deg.zero();
    LB_LAUNCH(c, incidence_fill4, cdiv(nt, 256 * kIncEpt), 256, 0, mesh->t4m.p, nt, mesh->k, aptr.p, deg.p, inc4.p);
    phase(c, "incidence");
    // clamped elements: the aniso numerators are fp64 even for fp32 meshes (solver.py:278-280)
    const bool degen_f32 = mesh->v_dtype == LB_F32 && kind != LB_FEM_TRIA_ANISO;
    if (mesh->k == 3) run_rows<3>(mesh, rec.p, consts.p, aptr.p, inc4.p, want_a, lump != 0, degen_f32, a_out, b_out);
    else run_rows<4>(mesh, rec.p, consts.p, aptr.p, inc4.p, want_a, lump != 0, degen_f32, a_out, b_out);
    tag();
    sync(c);  // u1/u2/aniso_mat are borrowed host buffers
    phase(c, "rows");
    LB_API_END
}

}  // extern "C"
