// kernels_sv_finish.cu - the tail of stage 9 (SV): varsv_sigh (draw.h:498-512: the squared increments are summed over ALL
// series - quirk draw.h:508 - and the shape uses the integer n / 2) and varsv_h0 (draw.h:523-537: dense k x k precision
// K0 = P0 + diag(1 / sigma_h^2), h0 ~ N(K0^-1 (P0 mu0 + h_1 / sigma_h^2), K0^-1)).
// One WARP per unit: k gamma draws, then the k x k system on the register-blocked warp Cholesky (warp_linalg.cuh) instead of
// the element-wise shared-memory one the first version ran inside a 128-thread CTA (0.09 ms at C3, 0.18 ms at C4 for a
// k x k system per unit: it sits at the end of the sweep's critical path).
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "warp_linalg.cuh"

namespace bvhar {

namespace {
constexpr int SF_MAXW = 4;

template <int NC>
__global__ void __launch_bounds__(32 * SF_MAXW) sv_finish_kernel(const Dims D, const DevPtrs P, const int wstride) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double colbuf[SF_MAXW][64];
  __shared__ int bad[SF_MAXW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int u = blockIdx.x * nw + warp;
  if (u >= D.U) return;
  const int w = u / D.C, c = u % D.C;
  const int k = D.k, n = P.win_n[w];
  const long long off = P.win_off[w] + (long long)c * k;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  double* Kp = sm + (size_t)warp * wstride;  // packed k(k+1)/2
  double* r = Kp + ((k * (k + 1) / 2 + 1) & ~1);
  double* z = r + k;
  double* x = z + k;
  double* sig = x + k;
  if (lane == 0) bad[warp] = 0;
  // S = sum over ALL series and t of (h_t - h_{t-1})^2 (quirk draw.h:508): the per-series sums come from sv_solve
  double ss = 0.0;
  for (int i = 0; i < k; ++i) ss += P.znorm[(size_t)D.U * k + (size_t)u * k + i];  // fixed order: deterministic
  for (int i = lane; i < k; i += 32) {
    const double s = 1.0 / rng.gamma(ST_SV_SIGH, (uint32_t)i, P.sig_shp[i] + (double)(n / 2), 1.0 / (P.sig_scl[i] + ss / 2));
    sig[i] = s;
    P.lvol_sig[(size_t)u * k + i] = s;
  }
  __syncwarp();
  for (int a = 0; a < k; ++a)
    for (int b = lane; b <= a; b += 32) Kp[a * (a + 1) / 2 + b] = P.sv_prec[a * k + b] + (a == b ? 1.0 / sig[a] : 0.0);
  for (int i = lane; i < k; i += 32) {
    double s = 0.0;
    for (int m = 0; m < k; ++m) s += P.sv_prec[i * k + m] * P.sv_mean[m];
    r[i] = s + P.H[off + i] / sig[i];
    z[i] = rng.normal_elem(ST_SV_H0_Z, (uint32_t)i);
  }
  __syncwarp();
  chol_blocked_warp<NC>(Kp, k, &bad[warp], colbuf[warp]);
  precision_solve_blocked<NC>(Kp, k, r, z, x);
  __syncwarp();
  for (int i = lane; i < k; i += 32) P.lvol_init[(size_t)u * k + i] = x[i];
  if (lane == 0 && bad[warp]) atomicOr(&P.flags[u], 4);
}

template <int NC>
cudaError_t launch_nc(const Dims& D, const DevPtrs& P, cudaStream_t st) {
  const int k = D.k;
  const int wstride = ((k * (k + 1) / 2 + 1) & ~1) + 4 * ((k + 1) & ~1);
  int nw = SF_MAXW;
  while (nw > 1 && (size_t)nw * wstride * sizeof(double) > 100 * 1024) nw >>= 1;
  const size_t smem = (size_t)nw * wstride * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaFuncSetAttribute(sv_finish_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sv_finish_kernel<NC><<<(D.U + nw - 1) / nw, nw * 32, smem, st>>>(D, P, wstride);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_sv_finish(const Dims& D, const DevPtrs& P, cudaStream_t st) {
  if (D.k <= 32) return launch_nc<1>(D, P, st);
  if (D.k <= 64) return launch_nc<2>(D, P, st);
  if (D.k <= 128) return launch_nc<4>(D, P, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace bvhar
