// dense.cu - small dense linear algebra for the Rayleigh-Ritz step (placeholder until lobpcg.cu lands)
#include "common.cuh"
namespace lb {
void destroy_dense_handles(lb_ctx *) {}
}  // namespace lb
