// kernels_pbuild.cu - stage 4, part A1 (the SV-weighted precision build) and the dispatcher of the small-design
// coefficient stage (parts: 0 = P build, 1 = batched Cholesky, 2 = sequential draws).
#include "coef_sizes.h"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"

namespace bvhar {

namespace {

// ---- A1: P_j = sum_{i>=j} L_ij^2 S_i, all j, one pass over the unit's S ----------------------
// Per unit this is a [k x k] x [k x Pp] product (upper-triangular weights W_ji = L_ij^2): DMMA m8n8k4 with the
// weights held in registers as A fragments for the whole kernel and the B fragments S_i[p] loaded straight
// from global memory (each S element is read exactly once, each P element written once: HBM-bound).
// MT = ceil(k / 8) row tiles (j), KS = ceil(k / 4) k-steps (i).
// With Qn != nullptr it also emits the cross matrices Q_j = sum_{i>=j+1} L_{i,j+1} L_ij S_i (slot j of Qn): Q_j delta_j is
// the only part of equation j+1's right-hand side that depends on equation j's draw (coef_seq_pipe_kernel).
template <int MT, int KS>
__global__ void __launch_bounds__(256) coef_pbuild_kernel(const Dims D, const DevPtrs P, double* __restrict__ Pfac, double* __restrict__ Qn) {
  const int u = blockIdx.x, k = D.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const double* Lu = P.L + (size_t)u * k * k;
  const int mtb = blockIdx.z * MT;  // first row tile of this pass (k > 32 takes two passes over S)
  double af[MT][KS];  // A[m = j][kk = i] = L_ij^2 for i >= j
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int j = 8 * (mtb + mt) + fr, i = 4 * ks + fc;
      const double l = (i < k && j <= i) ? Lu[i * k + j] : 0.0;
      af[mt][ks] = l * l;
    }
  double aq[MT][KS];  // A2[m = j][kk = i] = L_{i,j+1} L_ij for i >= j + 1
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int j = 8 * (mtb + mt) + fr, i = 4 * ks + fc;
      aq[mt][ks] = (Qn && i < k && j + 1 <= i) ? Lu[i * k + j] * Lu[i * k + j + 1] : 0.0;
    }
  double* Qu = Qn ? Qn + (size_t)u * k * D.ldS : nullptr;
  const double* Su = P.S + (size_t)u * k * D.ldS;
  double* Pu = Pfac + (size_t)u * k * D.ldS;
  const int ntiles = (D.Pp + 7) / 8;
  constexpr int NQ = KS > 8 ? 2 : 4;  // n-tiles per warp iteration (loads of all of them in flight together)
  for (int nt0 = (blockIdx.y * 8 + warp) * NQ; nt0 < ntiles; nt0 += gridDim.y * 8 * NQ) {
    double bf[NQ][KS];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int i = 4 * ks + fc, p = 8 * (nt0 + q) + fr;
        bf[q][ks] = (i < k && p < D.Pp) ? Su[(size_t)i * D.ldS + p] : 0.0;
      }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int p = 8 * (nt0 + q) + 2 * fc;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma_m8n8k4(c0, c1, af[mt][ks], bf[q][ks]);
        const int j = 8 * (mtb + mt) + fr;
        if (j < k && p < D.Pp) {  // ldS is even and p is even: 16-byte aligned pair inside the padded row
          *reinterpret_cast<double2*>(Pu + (size_t)j * D.ldS + p) = make_double2(c0, c1);
        }
        if (Qu) {  // warp-uniform
          double q0 = 0.0, q1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) dmma_m8n8k4(q0, q1, aq[mt][ks], bf[q][ks]);
          if (j + 1 < k && p < D.Pp) *reinterpret_cast<double2*>(Qu + (size_t)j * D.ldS + p) = make_double2(q0, q1);
        }
      }
    }
  }
}
// generic fallback for k > 64: one thread per (j, p), streaming i
__global__ void __launch_bounds__(256) coef_pbuild_generic_kernel(const Dims D, const DevPtrs P, double* __restrict__ Pfac) {
  const int u = blockIdx.x, k = D.k;
  const double* Su = P.S + (size_t)u * k * D.ldS;
  const double* Lu = P.L + (size_t)u * k * k;
  double* Pu = Pfac + (size_t)u * k * D.ldS;
  for (int j = 0; j < k; ++j)
    for (int p = threadIdx.x; p < D.Pp; p += blockDim.x) {
      double acc = 0.0;
      for (int i = j; i < k; ++i) {
        const double l = Lu[i * k + j];
        acc = fma(l * l, Su[(size_t)i * D.ldS + p], acc);
      }
      Pu[(size_t)j * D.ldS + p] = acc;
    }
}

}  // namespace

bool coef_v2_fits(const Dims& D, size_t smem_limit) {
  if (D.d > 128) return false;  // larger designs take the CTA-per-matrix path (kernels_chol_big.cu / kernels_seq_big.cu)
  const size_t ppad = (D.Pp + 1) & ~1;
  const size_t sm3 = coef_seq_smem_doubles<true>(D.k, D.d, D.ldS) * sizeof(double);
  return sm3 <= smem_limit && ppad * sizeof(double) <= smem_limit;
}

bool coef_pipe_fits(const Dims& D, size_t smem_limit) {
  // two CTAs per SM keep the chain latency hidden: use the pipelined kernel only while two of them fit
  return D.sv && D.k > 1 && D.k <= 64 && D.d <= 128 && 2 * coef_pipe_smem_doubles(D.k, D.d, D.ldS) * sizeof(double) <= smem_limit;
}

cudaError_t launch_coef_v2(int part, const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st) {
  const int nc = (D.d + 31) / 32, k = D.k;
  if (part == 0) {
    if (!D.sv) return cudaSuccess;
    const dim3 grid(D.U, 4);
    if (k <= 8) coef_pbuild_kernel<1, 2><<<grid, 256, 0, st>>>(D, P, Pfac, Qn);
    else if (k <= 16) coef_pbuild_kernel<2, 4><<<grid, 256, 0, st>>>(D, P, Pfac, Qn);
    else if (k <= 24) coef_pbuild_kernel<3, 6><<<grid, 256, 0, st>>>(D, P, Pfac, Qn);
    else if (k <= 32) coef_pbuild_kernel<4, 8><<<grid, 256, 0, st>>>(D, P, Pfac, Qn);
    else if (k <= 64) coef_pbuild_kernel<4, 16><<<dim3(D.U, 4, 2), 256, 0, st>>>(D, P, Pfac, Qn);
    else coef_pbuild_generic_kernel<<<D.U, 256, 0, st>>>(D, P, Pfac);
    return cudaGetLastError();
  }
  if (nc > 4) return cudaErrorInvalidConfiguration;
  if (part == 1) return launch_coef_chol(nc <= 1 ? 1 : nc <= 2 ? 2 : 4, D, P, Pfac, smem_limit, st);
  if (nc <= 1) return launch_coef_seq_nc1(D, P, Pfac, Qn, smem_limit, st);
  if (nc <= 2) return launch_coef_seq_nc2(D, P, Pfac, Qn, smem_limit, st);
  return launch_coef_seq_nc4(D, P, Pfac, Qn, smem_limit, st);
}

}  // namespace bvhar
