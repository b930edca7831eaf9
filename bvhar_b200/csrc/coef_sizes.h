// coef_sizes.h - shared-memory budgets of the small-design coefficient kernels (host + device).
#pragma once
#include <cstddef>

namespace bvhar {

constexpr int SEQ_THREADS = 256;
template <bool SV>
__host__ __device__ inline size_t coef_seq_smem_doubles(int k, int d, int ppad) {
  return 2 * (size_t)ppad + 3 * (size_t)k * d + (size_t)k * k + 4 * (size_t)d + (size_t)(SEQ_THREADS / 32) * d + (size_t)k + 8;
}

__host__ __device__ inline size_t coef_pipe_smem_doubles(int k, int d, int ppad) {
  return 4 * (size_t)ppad + 3 * (size_t)k * d + (size_t)k * k + 6 * (size_t)d + (size_t)(SEQ_THREADS / 32) * d + 8;
}

}  // namespace bvhar
