// block_linalg.cuh - small shared-memory dense kernels on packed-lower (row-major packed) matrices used by
// several sweep kernels: warp / block reductions, element-wise Cholesky by a block or a warp, and the
// "mean + chol^-T z" precision solve of draw_coef (bayes/misc/draw.h:372-383).
#pragma once
#include <cuda_runtime.h>

#include "engine.cuh"

namespace bvhar {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the block; every thread gets the result.  `scratch` holds >= 33 doubles.
__device__ inline double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// ------------------------------------------------------------------ packed Cholesky / solves
// Right-looking Cholesky of a packed-lower (row-major packed) matrix in shared memory, whole block.
__device__ inline void chol_packed_block(double* Pk, int d, int* bad) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int c = 0; c < d; ++c) {
    const int cc = tri_idx(c, c);
    __syncthreads();
    const double p = Pk[cc];
    __syncthreads();
    const double inv = rsqrt(p), piv = p * inv;  // one MUFU + Newton instead of sqrt followed by a division
    if (tid == 0) {
      Pk[cc] = piv;
      if (!(p > 0.0)) *bad = 1;
    }
    for (int a = c + 1 + tid; a < d; a += nt) Pk[tri_idx(a, c)] *= inv;
    __syncthreads();
    for (int a = c + 1 + warp; a < d; a += nw) {
      const double lac = Pk[tri_idx(a, c)];
      const int ra = a * (a + 1) / 2;
      for (int b = c + 1 + lane; b <= a; b += 32) Pk[ra + b] -= lac * Pk[tri_idx(b, c)];
    }
  }
  __syncthreads();
}
// Same for one warp (small matrices), packed-lower in shared or global memory.
__device__ inline void chol_packed_warp(double* Pk, int d, int* bad) {
  const int lane = threadIdx.x & 31;
  for (int c = 0; c < d; ++c) {
    const int cc = tri_idx(c, c);
    __syncwarp();
    const double p = Pk[cc];
    __syncwarp();
    const double inv = rsqrt(p), piv = p * inv;  // one MUFU + Newton instead of sqrt followed by a division
    if (lane == 0) {
      Pk[cc] = piv;
      if (!(p > 0.0)) *bad = 1;
    }
    for (int a = c + 1 + lane; a < d; a += 32) Pk[tri_idx(a, c)] *= inv;
    __syncwarp();
    for (int a = c + 1; a < d; ++a) {
      const double lac = Pk[tri_idx(a, c)];
      const int ra = a * (a + 1) / 2;
      for (int b = c + 1 + lane; b <= a; b += 32) Pk[ra + b] -= lac * Pk[tri_idx(b, c)];
    }
  }
  __syncwarp();
}
// One warp: x <- L^-T (L^-1 r + z), i.e. mean + chol(P)^-T z of draw.h:382-383.  x, r, z length d.
__device__ inline void precision_solve_warp(const double* Lp, int d, const double* r, const double* z, double* x) {
  const int lane = threadIdx.x & 31;
  for (int a = 0; a < d; ++a) {  // forward, row-oriented
    const int ra = a * (a + 1) / 2;
    double s = 0.0;
    for (int b = lane; b < a; b += 32) s += Lp[ra + b] * x[b];
    s = warp_sum(s);
    if (lane == 0) x[a] = (r[a] - s) / Lp[ra + a];
    __syncwarp();
  }
  for (int a = lane; a < d; a += 32) x[a] += z[a];
  __syncwarp();
  for (int c = d - 1; c >= 0; --c) {  // backward with L^T, column c of L
    double s = 0.0;
    for (int a = c + 1 + lane; a < d; a += 32) s += Lp[tri_idx(a, c)] * x[a];
    s = warp_sum(s);
    if (lane == 0) x[c] = (x[c] - s) / Lp[tri_idx(c, c)];
    __syncwarp();
  }
}

}  // namespace bvhar
