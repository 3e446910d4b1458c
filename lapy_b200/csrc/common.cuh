// common.cuh - context, error handling, device buffers and small device primitives shared by
// all translation units of liblapyb200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lapy_b200.h"
#include "exact.cuh"

namespace lb {

constexpr int kSMs = 148;          // B200: 2 dies x 74 SMs
constexpr double kEps = 2.220446049250313e-16;  // sys.float_info.epsilon (lapy/solver.py:158)

void set_error(const char *fmt, ...);

struct Error {
    int code;
};

#define LB_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            lb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,        \
                          __LINE__, cudaGetErrorString(_e));                                   \
            throw lb::Error{_e == cudaErrorMemoryAllocation ? LB_ERR_OOM : LB_ERR_CUDA};       \
        }                                                                                      \
    } while (0)

#define LB_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            lb::set_error(__VA_ARGS__);                                                        \
            throw lb::Error{LB_ERR_ARG};                                                       \
        }                                                                                      \
    } while (0)

// wraps an extern "C" body: C++ exceptions never cross the ABI
#define LB_API_BEGIN try {
#define LB_API_END                                                                             \
    }                                                                                          \
    catch (const lb::Error &e) { return e.code; }                                              \
    catch (const std::bad_alloc &) {                                                           \
        lb::set_error("host allocation failed");                                               \
        return LB_ERR_OOM;                                                                     \
    }                                                                                          \
    catch (...) {                                                                              \
        lb::set_error("unknown internal error");                                               \
        return LB_ERR_CUDA;                                                                    \
    }                                                                                          \
    return LB_OK;

}  // namespace lb

struct lb_dist;
struct lb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    int64_t n_strip_assemblies = 0, n_record_assemblies = 0;  // which assembly pipeline ran (lb_ctx_counters)
    void *pinned = nullptr;  // small pinned staging area for scalar read-backs
    size_t pinned_bytes = 0;
    void *cusolver = nullptr;  // cusolverDnHandle_t, created lazily (dense Rayleigh-Ritz only)
    void *cublas = nullptr;    // cublasHandle_t, created lazily (bring-up path of the dense block products)
    bool trace = false;        // LAPY_B200_TRACE=1: per-phase wall clock (synchronising!) on stderr
    double trace_t0 = 0;
    // per-kernel-class device timing (lb_profile_enable): event pairs around the hot launches
    int profile = 0;  // 0 off, 1 all classes, 2 SpMM class only (cheap enough for a timed region)
    struct ProfRec {
        int cls;
        cudaEvent_t e0, e1;
        double work;  // algorithmic bytes (HBM-bound classes) or flops (dense classes)
        int64_t shape[2];  // launch shape for the per-shape report: SpMM (columns, nnz), dense (p, q)
    };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
    // eigensolver work blocks ([X P W], A., B. double-buffered: 24 GB at level 9), kept between calls: the
    // stream-ordered pool re-maps physical memory to satisfy a fresh 4 GB request after the small
    // allocations of the multigrid setup split its free blocks (measured: 0.03 - 0.5 s per solve)
    void *ws = nullptr;
    size_t ws_bytes = 0;
    lb_dist *dist = nullptr;  // NCCL communicator of the row-partitioned mode (lb_comm_init)
    // two pinned staging buffers for large device -> pageable-host results (eigenvectors)
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};

namespace lb {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// stream-ordered device buffer (cudaMallocAsync pool of the context's device)
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    lb_ctx *ctx = nullptr;
    DBuf() = default;
    DBuf(lb_ctx *c, size_t count) { alloc(c, count); }
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), ctx(o.ctx) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; ctx = o.ctx;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    void alloc(lb_ctx *c, size_t count) {
        release();
        ctx = c;
        n = count;
        if (count == 0) { p = nullptr; return; }
        LB_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), c->stream));
    }
    void zero() {
        if (n) LB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream));
    }
    void release() {
        if (p) cudaFreeAsync(p, ctx->stream);
        p = nullptr;
        n = 0;
    }
    ~DBuf() { release(); }
    T *get() const { return p; }
    operator T *() const { return p; }
};

inline void h2d(lb_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes) LB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
}
inline void d2h(lb_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes) LB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
}
inline void d2d(lb_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes) LB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
}
inline void sync(lb_ctx *c) { LB_CUDA(cudaStreamSynchronize(c->stream)); }
// large result download into pageable host memory: chunked through pinned staging buffers, the
// host-side copy of chunk i (4 threads) overlaps the DMA of chunk i+1.  Synchronous.
void d2h_large(lb_ctx *c, void *dst, const void *src, size_t bytes);
// Touches every page of a large, freshly allocated host result buffer on helper threads WHILE the
// device computes, so that the final download copies into resident pages.  (A fresh 1 GB NumPy array
// is untouched anonymous memory: first-touch faults inside the download cost 0.1-0.5 s, box dependent -
// measured as the step-to-step spread of the ShapeDNA bench.)  Joined by the destructor / wait().
struct HostPrefault {
    std::vector<std::thread> th;
    HostPrefault(void *dst, size_t bytes, int nthreads = 2);
    void wait();
    ~HostPrefault() { wait(); }
};

// read back a few scalars (stream-ordered, then wait)
template <class T>
inline void read_back(lb_ctx *c, T *host, const T *dev, size_t count) {
    LB_REQUIRE(count * sizeof(T) <= c->pinned_bytes, "read_back too large");
    LB_CUDA(cudaMemcpyAsync(c->pinned, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    sync(c);
    std::memcpy(host, c->pinned, count * sizeof(T));
}

// the context's persistent workspace, grown on demand (contents undefined)
void *ctx_workspace(lb_ctx *c, size_t bytes);

double wall_ms();
// development aid: prints the wall time since the previous phase mark (after a stream sync)
inline void phase(lb_ctx *c, const char *name) {
    if (!c->trace) return;
    cudaStreamSynchronize(c->stream);
    double t = wall_ms();
    fprintf(stderr, "[lb trace] %-28s %9.3f ms\n", name, t - c->trace_t0);
    c->trace_t0 = wall_ms();
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// kernel launch with bookkeeping (gpu_launches in bench.py is this counter)
#define LB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                         \
    do {                                                                                       \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                       \
        (ctx)->launches++;                                                                     \
        LB_CUDA(cudaGetLastError());                                                           \
    } while (0)

void destroy_dense_handles(lb_ctx *ctx);  // dense.cu

// kernel classes of the profile report
enum ProfClass { PROF_SPMM = 0, PROF_GRAM, PROF_UPDATE, PROF_TRSM, PROF_DOTS, PROF_ELEMENTWISE, PROF_NCLASS };
struct ProfScope {
    lb_ctx *c;
    int idx = -1;
    ProfScope(lb_ctx *ctx, int cls, double work, int64_t shape0 = 0, int64_t shape1 = 0);
    ~ProfScope();
};

// exclusive prefix sum of int32 counts: out[0..n] (n+1 entries, out[n] = total)
void exclusive_scan_i32(lb_ctx *ctx, const int32_t *in, int32_t *out, int64_t n);

}  // namespace lb

// ---- device objects ---------------------------------------------------------------------
// Locality renumbering of the vertices (Morton order of a 128^3 cell grid), shared by a mesh and
// the matrices assembled on it and computed lazily by the first solver that wants it.
struct lb_order {
    lb_ctx *ctx = nullptr;
    std::shared_ptr<lb::DBuf<lb::D4>> v4;  // vertex positions (kept alive for the lazy computation)
    int64_t n = 0;                         // matrix dimension (max referenced vertex + 1)
    lb::DBuf<int32_t> order, inv;          // new -> old, old -> new; empty until ensure()
    bool ready = false;
};
namespace lb {
void ensure_order(lb_order &o);  // assembly.cu
}

struct lb_mesh {
    lb_ctx *ctx = nullptr;
    int64_t nv = 0, nt = 0;
    int k = 3;              // vertices per element
    int v_dtype = LB_F64;   // dtype of the caller's vertices: element math runs in it
    // ---- caller's numbering (gradient / divergence kernels, results returned per caller element)
    std::shared_ptr<lb::DBuf<lb::D4>> v4s;  // owner of v4 (shared with lb_order)
    lb::DBuf<lb::D4> &v4ref() { return *v4s; }
    lb::DBuf<float4> v4f;   // (nv) fp32 xyz + pad, only when v_dtype == LB_F32
    lb::DBuf<int4> t4;      // (nt) int32 x4, triangles padded with -1: one 16-byte load
    // vertex -> (element, corner) incidence in the caller's numbering, built on first use by the
    // divergence: code = element*4 + corner, ascending per vertex
    lb::DBuf<int32_t> inc_ptr;  // (nv+1)
    lb::DBuf<int32_t> inc;      // (k*nt)
    int64_t n_ref = 0;          // max referenced vertex + 1 (matrix dimension)
    bool has_inc = false;
    // ---- solver layout, built once at upload (lb_mesh_create): the assembly reads and writes in
    // the locality numbering `ord` (Morton order of the vertices), so that every gather of the
    // element / row kernels stays inside a few cache lines and the matrices come out in the
    // numbering the solvers iterate in
    std::shared_ptr<lb_order> ord;  // handed to the matrices assembled from this mesh
    lb::DBuf<lb::D4> v4m;           // (n_ref) vertices in the new numbering: v4m[r] = v4[order[r]]
    lb::DBuf<float4> v4fm;          // fp32 twin
    // (nt) elements with NEW vertex ids, sorted by their smallest new vertex id (ties by element id);
    // triangles carry the caller's element id in .w
    lb::DBuf<int4> t4m;
    lb::DBuf<int32_t> eorig;        // (nt) caller's element id of every sorted element (sort key of the tet rows)
    // triangles: strips of 128 consecutive rows for the strip-cooperative assembly (assembly.cu)
    lb::DBuf<int32_t> kptr;         // (n_ref + 1) first sorted element whose smallest vertex is >= v
    lb::DBuf<int32_t> hptr, hlist;  // per strip: the elements touching it whose smallest vertex lies in an earlier strip
    bool has_strips = false;
    // topological constants measured at upload for the single-pass strip assembly: stored entries of the
    // operators, rows with an element, and whether every row / strip qualifies for the fast path
    int64_t strip_nnz = -1, strip_nlump = -1;
    bool strip_fast = false;
};

struct lb_mat {
    lb_ctx *ctx = nullptr;
    int64_t n = 0, nnz = 0;     // n = number of rows
    int64_t ncols = -1;         // -1: square (n); prolongators / restrictors are rectangular CSR
    lb::DBuf<int32_t> indptr;   // (n+1)
    lb::DBuf<int32_t> indices;  // (nnz) sorted, unique per row
    lb::DBuf<double> data;      // (nnz)
    bool diagonal = false;      // every stored entry is on the diagonal (lumped mass / identity)
    // Matrices assembled from a mesh are STORED in the mesh's locality numbering (Morton order of
    // the vertices, `ord`): row / column r of the stored matrix is vertex ord->order[r] of the caller.
    // The solvers iterate in that numbering; lb_mat_download, lb_spmm and the solver outputs convert
    // to the caller's numbering.  Uploaded matrices (lb_mat_upload) are in the caller's numbering
    // (permuted == false, ord == nullptr).
    bool permuted = false;
    std::shared_ptr<lb_order> ord;
    // caches of the strip-staged SpMM (blockvec.cu), filled on first use: the largest number of entries
    // in any strip of 8 / 16 / 32 / 64 / 128 consecutive rows, and a single-precision copy of the values
    mutable int32_t strip_cap[5] = {0, 0, 0, 0, 0};
    mutable bool strip_ready = false;
    mutable lb::DBuf<float> data32;
};
