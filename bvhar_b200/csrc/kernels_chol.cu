// kernels_chol.cu - stage 4, part A2: chol(diag(prec_j) + P_j) for every (unit, equation) of the batch, one
// warp per matrix, 32 x 32 register blocks (warp_linalg.cuh).  Compiled once per BV_NC = ceil(d / 32) in
// {1, 2, 4, 8} (separate translation units: the fully unrolled register-block code is slow to compile).
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"
#include "warp_linalg.cuh"

#ifndef BV_NC
#error "compile with -DBV_NC=1|2|4|8"
#endif
#define BV_CAT2(a, b) a##b
#define BV_CAT(a, b) BV_CAT2(a, b)

namespace bvhar {

namespace {
constexpr int NC = BV_NC;

// ---- A2: chol(diag(prec_j) + P_j), one warp per (unit, equation) ------------------------------
__global__ void __launch_bounds__(128) BV_CAT(coef_chol_kernel_nc, BV_NC)(const Dims D, const DevPtrs P, double* __restrict__ Pfac) {
  const bool SV = D.sv != 0;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long job = (long long)blockIdx.x * nw + warp;
  if (job >= (long long)D.U * D.k) return;
  const int u = (int)(job / D.k), j = (int)(job % D.k), w = u / D.C;
  const int k = D.k, d = D.d, Pp = D.Pp;
  double* Pk = sm + (size_t)warp * ((Pp + 1) & ~1);
  double* Pg = Pfac + ((size_t)u * k + j) * D.ldS;
  __shared__ int bad[8];
  __shared__ double colbuf[8][64];  // column broadcast buffers of the warp Cholesky (warp_linalg.cuh)
  if (lane == 0) bad[warp] = 0;
  if (SV) {  // 16-byte asynchronous copies: all of a lane's ~30 chunks in flight (both bases and the padded length are 16-byte multiples)
    const int ppad = (Pp + 1) & ~1;
    for (int p = 2 * lane; p < ppad; p += 64) cp_async16(Pk + p, Pg + p, 16);
    cp_async_commit();
    cp_async_wait<0>();
  } else {  // P_j = (sum_{i>=j} L_ij^2 / d_i) X'X
    double ws = 0.0;
    for (int i = j; i < k; ++i) {
      const double l = P.L[((size_t)u * k + i) * k + j];
      ws += l * l / P.diag[(size_t)u * k + i];
    }
    const double* X2 = P.XtX + (size_t)w * D.ldS;
    for (int p = lane; p < Pp; p += 32) Pk[p] = ws * X2[p];
  }
  __syncwarp();
  const double* pr = P.prec + (size_t)u * k * d + (size_t)j * d;
  for (int a = lane; a < d; a += 32) Pk[a * (a + 1) / 2 + a] += pr[a];
  __syncwarp();
  chol_blocked_warp<NC>(Pk, d, &bad[warp], colbuf[warp]);
  for (int p = lane; p < Pp; p += 32) Pg[p] = Pk[p];
  __syncwarp();
  if (lane == 0 && bad[warp]) atomicOr(&P.flags[u], 1);
}

}  // namespace

cudaError_t BV_CAT(launch_coef_chol_nc, BV_NC)(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st) {
  const size_t ppad = (D.Pp + 1) & ~1;
  int nw = 4;
  while (nw > 1 && nw * ppad * sizeof(double) > 96 * 1024) nw >>= 1;
  const size_t sm2 = nw * ppad * sizeof(double);
  if (sm2 > smem_limit) return cudaErrorInvalidConfiguration;
  const long long jobs = (long long)D.U * D.k;
  cudaFuncSetAttribute(BV_CAT(coef_chol_kernel_nc, BV_NC), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  BV_CAT(coef_chol_kernel_nc, BV_NC)<<<(unsigned)((jobs + nw - 1) / nw), nw * 32, sm2, st>>>(D, P, Pfac);
  return cudaGetLastError();
}

}  // namespace bvhar
