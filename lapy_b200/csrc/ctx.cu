// ctx.cu - context life cycle, error reporting, prefix sums, matrix upload / download.
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

#include "amg.cuh"

namespace lb {

static thread_local char g_err[1024] = "";

double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- large downloads -----------------------------------------------------------------------------
constexpr size_t kStageBytes = 64ull << 20;

static void parallel_memcpy(char *dst, const char *src, size_t bytes) {
    const int nt = 4;
    if (bytes < (8u << 20)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::thread th[nt - 1];
    const size_t part = (bytes / nt + 4095) & ~size_t(4095);
    for (int t = 1; t < nt; t++) {
        const size_t off = std::min(bytes, part * t), len = std::min(part, bytes - off);
        th[t - 1] = std::thread([=] { if (len) std::memcpy(dst + off, src + off, len); });
    }
    std::memcpy(dst, src, std::min(part, bytes));
    for (int t = 1; t < nt; t++) th[t - 1].join();
}

// The destination of a large download is usually a fresh NumPy allocation (anonymous mmap, untouched):
// filling 1 GB through 4 KB first-touch faults costs 0.1-0.5 s and varies from box to box.  Ask for
// transparent huge pages on the 2 MB-aligned interior (no-op where THP is off; errors ignored).
static void hint_huge_pages(void *dst, size_t bytes) {
    const uintptr_t huge = 2u << 20;
    const uintptr_t b = (reinterpret_cast<uintptr_t>(dst) + huge - 1) & ~(huge - 1);
    const uintptr_t e = (reinterpret_cast<uintptr_t>(dst) + bytes) & ~(huge - 1);
#ifdef MADV_HUGEPAGE
    if (e > b) (void)madvise(reinterpret_cast<void *>(b), e - b, MADV_HUGEPAGE);
#endif
}

static bool is_page_locked(const void *p) {
    cudaPointerAttributes attr{};
    const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // an unregistered pointer is not an error here
    return yes;
}

HostPrefault::HostPrefault(void *dst, size_t bytes, int nthreads) {
    if (!dst || bytes < (32u << 20) || is_page_locked(dst)) return;
    hint_huge_pages(dst, bytes);
    const size_t part = ((bytes / nthreads) + 4095) & ~size_t(4095);
    for (int t = 0; t < nthreads; t++) {
        const size_t off = std::min(bytes, part * t), len = std::min(part, bytes - off);
        if (len == 0) continue;
        th.emplace_back([=] {
            volatile char *p = static_cast<volatile char *>(dst) + off;
            for (size_t i = 0; i < len; i += 4096) p[i] = 0;  // the buffer is an output: its content is overwritten later
            p[len - 1] = 0;
        });
    }
}

void HostPrefault::wait() {
    for (auto &t : th)
        if (t.joinable()) t.join();
    th.clear();
}

void d2h_large(lb_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes < (32u << 20)) {
        d2h(c, dst, src, bytes);
        sync(c);
        return;
    }
    if (is_page_locked(dst)) {  // cudaHostAlloc / cudaHostRegister (lb_host_alloc, a pinned torch tensor): one DMA, no staging
        d2h(c, dst, src, bytes);
        sync(c);
        return;
    }
    hint_huge_pages(dst, bytes);
    for (int i = 0; i < 2; i++) {
        if (!c->stage[i]) {
            LB_CUDA(cudaMallocHost(&c->stage[i], kStageBytes));
            LB_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
        }
    }
    const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
    auto issue = [&](size_t k) {
        const size_t off = k * kStageBytes, len = std::min(kStageBytes, bytes - off);
        LB_CUDA(cudaMemcpyAsync(c->stage[k & 1], (const char *)src + off, len, cudaMemcpyDeviceToHost, c->stream));
        LB_CUDA(cudaEventRecord(c->stage_ev[k & 1], c->stream));
    };
    issue(0);
    for (size_t k = 0; k < nchunks; k++) {
        LB_CUDA(cudaEventSynchronize(c->stage_ev[k & 1]));
        if (k + 1 < nchunks) issue(k + 1);  // uses the other buffer, already drained
        const size_t off = k * kStageBytes, len = std::min(kStageBytes, bytes - off);
        parallel_memcpy((char *)dst + off, (const char *)c->stage[k & 1], len);
    }
}

// ---- per-class device timing ------------------------------------------------------------------
static cudaEvent_t prof_event(lb_ctx *c) {
    if (!c->prof_pool.empty()) {
        cudaEvent_t e = c->prof_pool.back();
        c->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void *ctx_workspace(lb_ctx *c, size_t bytes) {
    if (bytes > c->ws_bytes) {
        if (c->ws) LB_CUDA(cudaFreeAsync(c->ws, c->stream));
        c->ws = nullptr;
        c->ws_bytes = 0;
        LB_CUDA(cudaMallocAsync(&c->ws, bytes, c->stream));
        c->ws_bytes = bytes;
    }
    return c->ws;
}

ProfScope::ProfScope(lb_ctx *ctx, int cls, double work, int64_t shape0, int64_t shape1) : c(ctx) {
    if (c->profile == 0 || (c->profile == 2 && cls != PROF_SPMM)) return;
    lb_ctx::ProfRec r{cls, prof_event(c), prof_event(c), work, {shape0, shape1}};
    cudaEventRecord(r.e0, c->stream);
    idx = (int)c->prof.size();
    c->prof.push_back(r);
}

ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->prof[idx].e1, c->stream);
}

// ---- exclusive scan (int32) ---------------------------------------------------------------
// Three small kernels (tile reduce -> scan of tile sums -> tile scan + offset).  Inputs are at
// most a few tens of MB (per-vertex / per-row counts), so this is launch-latency, not
// bandwidth dominated; every pass is fully coalesced.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int x) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    return x;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int x, int *total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(x);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
        s = warp_incl_scan(s);
        warp_sums[lane] = s;
    }
    __syncthreads();
    int base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[((blockDim.x + 31) >> 5) - 1];
    __syncthreads();
    return base + incl - x;
}

__global__ void scan_tile_reduce(const int32_t *__restrict__ in, int32_t *__restrict__ tile_sums,
                                 int64_t n) {
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    int total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_tile_sums(int32_t *tile_sums, int ntiles, int32_t *grand_total) {
    // single block, loops over the tile sums in chunks of blockDim.x
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += blockDim.x) {
        int i = base + threadIdx.x;
        int x = i < ntiles ? tile_sums[i] : 0;
        int total;
        int ex = block_excl_scan(x, &total);
        int c = carry;
        if (i < ntiles) tile_sums[i] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void scan_tile_apply(const int32_t *__restrict__ in, const int32_t *__restrict__ tile_sums,
                                int32_t *__restrict__ out, int64_t n) {
    // each thread owns kScanItems CONSECUTIVE items (blocked arrangement) staged through smem
    __shared__ int32_t buf[kScanTile];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int l = i * kScanThreads + threadIdx.x;
        int64_t idx = base + l;
        buf[l] = idx < n ? in[idx] : 0;
    }
    __syncthreads();
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = buf[threadIdx.x * kScanItems + i];
        s += v[i];
    }
    int total;
    int ex = block_excl_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        buf[threadIdx.x * kScanItems + i] = ex;
        ex += v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int l = i * kScanThreads + threadIdx.x;
        int64_t idx = base + l;
        if (idx < n) out[idx] = buf[l];
    }
}

void exclusive_scan_i32(lb_ctx *ctx, const int32_t *in, int32_t *out, int64_t n) {
    if (n == 0) {
        LB_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
        return;
    }
    int ntiles = cdiv(n, kScanTile);
    DBuf<int32_t> tile_sums(ctx, ntiles);
    LB_LAUNCH(ctx, scan_tile_reduce, ntiles, kScanThreads, 0, in, tile_sums.p, n);
    LB_LAUNCH(ctx, scan_tile_sums, 1, 1024, 0, tile_sums.p, ntiles, out + n);
    LB_LAUNCH(ctx, scan_tile_apply, ntiles, kScanThreads, 0, in, tile_sums.p, out, n);
}

}  // namespace lb

using namespace lb;

extern "C" {

const char *lb_last_error(void) { return g_err; }
const char *lb_version(void) { return "lapy_b200 0.1 (sm_100a)"; }

int lb_ctx_create(int device, void *stream, lb_ctx **out) {
    LB_API_BEGIN
    LB_REQUIRE(out != nullptr, "lb_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("lb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return LB_ERR_CUDA;
    }
    LB_REQUIRE(device >= 0 && device < count, "lb_ctx_create: device %d out of range [0,%d)", device, count);
    DeviceGuard g(device);
    lb_ctx *c = new lb_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        LB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    LB_CUDA(cudaEventCreate(&c->ev0));
    LB_CUDA(cudaEventCreate(&c->ev1));
    c->pinned_bytes = 1 << 16;
    LB_CUDA(cudaMallocHost(&c->pinned, c->pinned_bytes));
    // keep freed blocks in the pool: assembly / solver workspaces are re-used across calls
    cudaMemPool_t pool;
    LB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    LB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    const char *tr = getenv("LAPY_B200_TRACE");
    c->trace = tr && tr[0] == '1';
    *out = c;
    LB_API_END
}

int lb_host_alloc(size_t bytes, void **out) {
    LB_API_BEGIN
    LB_REQUIRE(out && bytes > 0, "lb_host_alloc: bad argument");
    *out = nullptr;
    LB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    LB_API_END
}

int lb_host_free(void *p) {
    LB_API_BEGIN
    if (p) cudaFreeHost(p);
    cudaGetLastError();  // at interpreter exit the driver may already be gone
    LB_API_END
}

int lb_ctx_release_workspace(lb_ctx *c) {
    LB_API_BEGIN
    LB_REQUIRE(c, "ctx is NULL");
    DeviceGuard g(c->device);
    if (c->ws) LB_CUDA(cudaFreeAsync(c->ws, c->stream));
    c->ws = nullptr;
    c->ws_bytes = 0;
    LB_API_END
}

int lb_ctx_destroy(lb_ctx *c) {
    LB_API_BEGIN
    if (!c) return LB_OK;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    lb::destroy_dense_handles(c);
    for (auto &r : c->prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    if (c->ws) cudaFreeAsync(c->ws, c->stream);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaFreeHost(c->pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    LB_API_END
}

int lb_ctx_sync(lb_ctx *c) {
    LB_API_BEGIN
    LB_REQUIRE(c, "ctx is NULL");
    DeviceGuard g(c->device);
    sync(c);
    LB_API_END
}

int lb_timer_start(lb_ctx *c) {
    LB_API_BEGIN
    LB_REQUIRE(c, "ctx is NULL");
    DeviceGuard g(c->device);
    LB_CUDA(cudaEventRecord(c->ev0, c->stream));
    LB_API_END
}

int lb_timer_stop(lb_ctx *c, double *ms) {
    LB_API_BEGIN
    LB_REQUIRE(c && ms, "ctx/ms is NULL");
    DeviceGuard g(c->device);
    LB_CUDA(cudaEventRecord(c->ev1, c->stream));
    LB_CUDA(cudaEventSynchronize(c->ev1));
    float f = 0;
    LB_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    LB_API_END
}

int lb_profile_enable(lb_ctx *c, int on) {
    LB_API_BEGIN
    LB_REQUIRE(c, "ctx is NULL");
    DeviceGuard g(c->device);
    sync(c);
    for (auto &r : c->prof) {
        c->prof_pool.push_back(r.e0);
        c->prof_pool.push_back(r.e1);
    }
    c->prof.clear();
    c->profile = on;
    LB_API_END
}

// per class: launches, total device ms, total work (bytes or flops); arrays of PROF_NCLASS = 6
int lb_profile_report(lb_ctx *c, int64_t *count, double *ms, double *work) {
    LB_API_BEGIN
    LB_REQUIRE(c && count && ms && work, "NULL argument");
    DeviceGuard g(c->device);
    sync(c);
    for (int k = 0; k < PROF_NCLASS; k++) count[k] = 0, ms[k] = 0, work[k] = 0;
    for (auto &r : c->prof) {
        float t = 0;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        count[r.cls]++;
        ms[r.cls] += t;
        work[r.cls] += r.work;
    }
    LB_API_END
}

// records of one class aggregated by launch shape, largest device time first; arrays of `cap`
int lb_profile_shapes(lb_ctx *c, int cls, int cap, int64_t *shape0, int64_t *shape1, int64_t *count, double *ms,
                      double *work, int *nshapes) {
    LB_API_BEGIN
    LB_REQUIRE(c && shape0 && shape1 && count && ms && work && nshapes && cap > 0, "NULL argument");
    LB_REQUIRE(cls >= 0 && cls < PROF_NCLASS, "unknown profile class %d", cls);
    DeviceGuard g(c->device);
    sync(c);
    struct Agg {
        int64_t s0, s1, cnt;
        double ms, work;
    };
    std::vector<Agg> agg;
    for (auto &r : c->prof) {
        if (r.cls != cls) continue;
        float t = 0;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        Agg *a = nullptr;
        for (auto &x : agg)
            if (x.s0 == r.shape[0] && x.s1 == r.shape[1]) a = &x;
        if (!a) {
            agg.push_back({r.shape[0], r.shape[1], 0, 0.0, 0.0});
            a = &agg.back();
        }
        a->cnt++;
        a->ms += t;
        a->work += r.work;
    }
    std::sort(agg.begin(), agg.end(), [](const Agg &x, const Agg &y) { return x.ms > y.ms; });
    const int k = (int)std::min<size_t>(agg.size(), (size_t)cap);
    for (int i = 0; i < k; i++) {
        shape0[i] = agg[i].s0;
        shape1[i] = agg[i].s1;
        count[i] = agg[i].cnt;
        ms[i] = agg[i].ms;
        work[i] = agg[i].work;
    }
    *nshapes = k;
    LB_API_END
}

int lb_launch_count(lb_ctx *c, int64_t *count) {
    LB_API_BEGIN
    LB_REQUIRE(c && count, "ctx/count is NULL");
    *count = c->launches;
    LB_API_END
}

// [0] kernels launched, [1] assemblies done by the strip-cooperative kernels, [2] assemblies done by
// the record pipeline (tets, and triangle meshes the strip kernels do not cover), [3] reserved
int lb_ctx_counters(lb_ctx *c, int64_t *out4) {
    LB_API_BEGIN
    LB_REQUIRE(c && out4, "ctx/out is NULL");
    out4[0] = c->launches;
    out4[1] = c->n_strip_assemblies;
    out4[2] = c->n_record_assemblies;
    out4[3] = 0;
    LB_API_END
}

// ---- matrices ---------------------------------------------------------------------------
int lb_mat_info(lb_mat *m, int64_t *n, int64_t *nnz) {
    LB_API_BEGIN
    LB_REQUIRE(m, "matrix is NULL");
    if (n) *n = m->n;
    if (nnz) *nnz = m->nnz;
    LB_API_END
}

int lb_mat_download(lb_mat *m, int32_t *indptr, int32_t *indices, double *data) {
    LB_API_BEGIN
    LB_REQUIRE(m, "matrix is NULL");
    lb_ctx *c = m->ctx;
    DeviceGuard g(c->device);
    // matrices assembled on the device live in the mesh's locality numbering: the caller gets the
    // canonical CSC of ITS numbering (rows / columns permuted back, rows sorted)
    std::unique_ptr<lb_mat> plain;
    if (m->permuted) {
        plain = to_caller_order(c, m);
        m = plain.get();
    }
    if (indptr) d2h_large(c, indptr, m->indptr.p, (m->n + 1) * sizeof(int32_t));
    if (indices) d2h_large(c, indices, m->indices.p, m->nnz * sizeof(int32_t));
    if (data) d2h_large(c, data, m->data.p, m->nnz * sizeof(double));
    sync(c);
    LB_API_END
}

__global__ void check_diagonal_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                      int64_t n, int *not_diag) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int p = indptr[r]; p < indptr[r + 1]; p++)
        if (indices[p] != r) *not_diag = 1;
}

int lb_mat_upload(lb_ctx *c, int64_t n, int64_t nnz, const int32_t *indptr, const int32_t *indices,
                  const double *data, lb_mat **out) {
    LB_API_BEGIN
    LB_REQUIRE(c && out, "ctx/out is NULL");
    LB_REQUIRE(n >= 0 && nnz >= 0 && n < INT32_MAX && nnz < INT32_MAX, "matrix too large for int32 indices");
    LB_REQUIRE(indptr && (nnz == 0 || (indices && data)), "NULL array");
    LB_REQUIRE(indptr[0] == 0 && indptr[n] == nnz, "indptr does not match nnz");
    DeviceGuard g(c->device);
    lb_mat *m = new lb_mat();
    m->ctx = c;
    m->n = n;
    m->nnz = nnz;
    try {
        m->indptr.alloc(c, n + 1);
        m->indices.alloc(c, nnz);
        m->data.alloc(c, nnz);
        h2d(c, m->indptr.p, indptr, (n + 1) * sizeof(int32_t));
        h2d(c, m->indices.p, indices, nnz * sizeof(int32_t));
        h2d(c, m->data.p, data, nnz * sizeof(double));
        DBuf<int> flag(c, 1);
        flag.zero();
        if (n) LB_LAUNCH(c, check_diagonal_kernel, cdiv(n, 256), 256, 0, m->indptr.p, m->indices.p, n, flag.p);
        int h = 0;
        read_back(c, &h, flag.p, 1);
        m->diagonal = (h == 0);
    } catch (...) {
        delete m;
        throw;
    }
    *out = m;
    LB_API_END
}

int lb_mat_free(lb_mat *m) {
    LB_API_BEGIN
    if (!m) return LB_OK;
    DeviceGuard g(m->ctx->device);
    delete m;
    LB_API_END
}

}  // extern "C"
