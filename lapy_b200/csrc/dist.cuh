// dist.cuh - row-partitioned (multi-GPU, NCCL) helpers; see dist.cu
#pragma once
#include "blockvec.cuh"

struct lb_dist {
    void *comm = nullptr;  // ncclComm_t
    int rank = 0, world = 1;
};

namespace lb {
using DistCtx = lb_dist;

// in-place sum over ranks of a small device buffer (Gram matrices, column dots)
void dist_allreduce_sum(lb_ctx *c, const DistCtx *d, double *buf, size_t count);
// recv (world * count_per_rank) = concatenation over ranks of send (count_per_rank)
void dist_allgather(lb_ctx *c, const DistCtx *d, const double *send, double *recv, size_t count_per_rank);
void dist_allgather_i32(lb_ctx *c, const DistCtx *d, const int32_t *send, int32_t *recv, size_t count_per_rank);
// grouped point-to-point exchange (halo): for every peer p != rank, send cnt_s[p] elements starting
// at send + off_s[p] and receive cnt_r[p] elements into recv + off_r[p]; elem_size 4 (int32) or 8 (fp64)
void dist_exchange(lb_ctx *c, const DistCtx *d, const void *send, const int64_t *off_s, const int64_t *cnt_s,
                   void *recv, const int64_t *off_r, const int64_t *cnt_r, int elem_size);

// plain cudaMalloc buffer: NCCL transports (P2P / IPC) must not be handed stream-ordered pool memory
template <class T>
struct RawBufT {
    T *p = nullptr;
    size_t n = 0;
    RawBufT() = default;
    RawBufT(const RawBufT &) = delete;
    RawBufT &operator=(const RawBufT &) = delete;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) LB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~RawBufT() { release(); }
};
}  // namespace lb
