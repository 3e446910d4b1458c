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
}  // namespace lb
