// dist.cu - NCCL plumbing for the row-partitioned single-mesh mode: one process
// per GPU, contiguous row blocks of the renumbered operator per rank, NCCL all-gather of the block
// vectors before every SpMM (first form of the halo exchange) and NCCL all-reduce for the Gram
// matrices and column dots.  The reference has no distributed mode at all (single process).
//
// NCCL is loaded lazily with dlopen so that single-GPU users never touch it; inside a torchrun job
// the already loaded libnccl.so.2 of PyTorch is picked up (same soname).
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is dlopen()ed, never linked

#include "dist.cuh"

namespace lb {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
constexpr ncclResult_t ncclSuccessV = ncclSuccess;
constexpr ncclDataType_t ncclFloat64V = ncclDouble;
constexpr ncclRedOp_t ncclSumV = ncclSum;

static NcclApi &nccl() {
    static NcclApi api;
    if (!api.handle) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            set_error("cannot load libnccl.so.2: %s", dlerror());
            throw Error{LB_ERR_UNSUPPORTED};
        }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
        api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather) {
            set_error("libnccl.so.2 lacks the expected symbols");
            throw Error{LB_ERR_UNSUPPORTED};
        }
        api.handle = h;
    }
    return api;
}

#define LB_NCCL(expr)                                                                          \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccessV) {                                                              \
            set_error("NCCL error %d at %s:%d: %s", (int)_r, __FILE__, __LINE__,                    \
                      nccl().GetErrorString ? nccl().GetErrorString(_r) : "?");                \
            throw Error{LB_ERR_CUDA};                                                          \
        }                                                                                      \
    } while (0)

void dist_allreduce_sum(lb_ctx *c, const DistCtx *d, double *buf, size_t count) {
    if (!d || d->world == 1 || count == 0) return;
    LB_NCCL(nccl().AllReduce(buf, buf, count, ncclFloat64V, ncclSumV, (ncclComm_t)d->comm, c->stream));
}

void dist_allgather(lb_ctx *c, const DistCtx *d, const double *send, double *recv, size_t count_per_rank) {
    LB_NCCL(nccl().AllGather(send, recv, count_per_rank, ncclFloat64V, (ncclComm_t)d->comm, c->stream));
}

void dist_allgather_i32(lb_ctx *c, const DistCtx *d, const int32_t *send, int32_t *recv, size_t count_per_rank) {
    LB_NCCL(nccl().AllGather(send, recv, count_per_rank, ncclInt32, (ncclComm_t)d->comm, c->stream));
}

void dist_exchange(lb_ctx *c, const DistCtx *d, const void *send, const int64_t *off_s, const int64_t *cnt_s,
                   void *recv, const int64_t *off_r, const int64_t *cnt_r, int elem_size) {
    NcclApi &api = nccl();
    if (!api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd) {
        set_error("libnccl.so.2 lacks ncclSend/ncclRecv (needs NCCL >= 2.7)");
        throw Error{LB_ERR_UNSUPPORTED};
    }
    const ncclDataType_t dt = elem_size == 8 ? ncclDouble : ncclInt32;
    const char *sp = static_cast<const char *>(send);
    char *rp = static_cast<char *>(recv);
    LB_NCCL(api.GroupStart());
    for (int p = 0; p < d->world; p++) {
        if (p == d->rank) continue;
        if (cnt_s[p] > 0)
            LB_NCCL(api.Send(sp + (size_t)off_s[p] * elem_size, (size_t)cnt_s[p], dt, p, (ncclComm_t)d->comm, c->stream));
        if (cnt_r[p] > 0)
            LB_NCCL(api.Recv(rp + (size_t)off_r[p] * elem_size, (size_t)cnt_r[p], dt, p, (ncclComm_t)d->comm, c->stream));
    }
    LB_NCCL(api.GroupEnd());
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_nccl_unique_id(unsigned char *out128) {
    LB_API_BEGIN
    LB_REQUIRE(out128, "out is NULL");
    ncclUniqueId id;
    LB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out128, id.internal, 128);
    LB_API_END
}

int lb_comm_init(lb_ctx *c, int world, int rank, const unsigned char *id128) {
    LB_API_BEGIN
    LB_REQUIRE(c && id128 && world >= 1 && rank >= 0 && rank < world, "lb_comm_init: bad argument");
    DeviceGuard g(c->device);
    if (c->dist) {
        if (c->dist->comm) nccl().CommDestroy((ncclComm_t)c->dist->comm);
        delete c->dist;
        c->dist = nullptr;
    }
    ncclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    LB_NCCL(nccl().CommInitRank(&comm, world, id, rank));
    c->dist = new DistCtx();
    c->dist->comm = comm;
    c->dist->rank = rank;
    c->dist->world = world;
    LB_API_END
}

int lb_comm_destroy(lb_ctx *c) {
    LB_API_BEGIN
    LB_REQUIRE(c, "ctx is NULL");
    if (c->dist) {
        DeviceGuard g(c->device);
        sync(c);
        if (c->dist->comm) nccl().CommDestroy((ncclComm_t)c->dist->comm);
        delete c->dist;
        c->dist = nullptr;
    }
    LB_API_END
}

}  // extern "C"
