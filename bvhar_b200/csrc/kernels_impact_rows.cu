// kernels_impact_rows.cu - stage 7 (+8), phase 2: the row draws of updateImpact (triangular.h:322-336 + draw_coef
// draw.h:372-383 + draw_savs draw.h:398-415) and the rebuilt rows of L (updateChol, triangular.h:348, build_inv_lower
// draw.h:124-132), ONE WARP PER (unit, row j) over the whole batch.
//
// Row j regresses Z_j / s_j on Z_{<j} / s_j: a j x j system whose Gram and right-hand side were left in the per-unit
// scratch by phase 1 (SV: impact_sv_kernel, packed per row; LDLT: the unit's full Z'Z).  The rows are conditionally
// independent, so all U (k - 1) of them run at once instead of k / 8 at a time inside the unit's CTA, and the factorisation
// is the register-blocked warp Cholesky of the coefficient stage (warp_linalg.cuh: 32 x 32 blocks in registers, blocked
// two-sided solve) instead of an element-wise shared-memory one.  Compiled once per BV_NC = ceil(j / 32) in {1, 2, 4};
// a launch covers the rows jlo .. jhi of every unit.
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "warp_linalg.cuh"

#ifndef BV_NC
#error "compile with -DBV_NC=1|2|4"
#endif
#define BV_CAT2(a, b) a##b
#define BV_CAT(a, b) BV_CAT2(a, b)

namespace bvhar {

namespace {
constexpr int NC = BV_NC;
constexpr int IR_MAXW = 4;

__device__ __forceinline__ int rows_row_off(int j) {  // sum_{j'<j} (j'(j'+1)/2 + j'), the packed layout of impact_sv_kernel
  return ((j - 1) * j * (2 * j - 1) / 6 + 3 * (j - 1) * j / 2) / 2;
}

#if BV_NC == 1
// j <= 32: the whole system is one register block (lane = row).  Same arithmetic as chol_blocked_warp<1> +
// precision_solve_blocked<1>, but every 32-step chain stops at step j (warp-uniform exit: the unrolled bodies keep their
// static register indices) - the rows of a k = 20 model average 10 unknowns, a third of the block.
__device__ __forceinline__ void small_system_solve(double* Pw, const int j, const double* r, const double* z, double* x, double* colbuf, int* bad) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool va = lane < j;
  const int ra = lane * (lane + 1) / 2;
  double dg[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) dg[c] = (va && c <= lane) ? Pw[ra + c] : (c == lane ? 1.0 : 0.0);
  // Cholesky, pivot chain without the shared-memory round trip (chol32_smem_chain)
  double dinv = 1.0;
  double pc = __shfl_sync(FULL, dg[0], 0);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c >= j) break;
    const double inv = rsqrt(pc), piv = pc * inv;
    if (lane == c && !(pc > 0.0)) *bad = 1;
    const double lc = lane > c ? dg[c] * inv : 0.0;
    if (lane == c) { dg[c] = piv; dinv = inv; }
    else dg[c] = lc;
    if (c < 31) pc = __shfl_sync(FULL, fma(-lc, lc, dg[c + 1]), c + 1);
    double* cb = colbuf + 32 * (c & 1);
    cb[lane] = lc;
    __syncwarp();
#pragma unroll
    for (int b = c + 1; b < 32; ++b) dg[b] = fma(-lc, cb[b], dg[b]);
  }
  __syncwarp();
  // the factor back to shared memory (the backward solve reads columns)
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (va && c <= lane) Pw[ra + c] = dg[c];
  __syncwarp();
  // forward: L y = r, lanes own rows, unknowns scaled by 1 / L_aa
  double y = va ? r[lane] * dinv : 0.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c >= j) break;
    const double yc = __shfl_sync(FULL, y, c);
    if (lane > c) y = fma(-(dg[c] * dinv), yc, y);
  }
  // backward: L' x = y + z, lanes own columns
  double xv = va ? (y + z[lane]) * dinv : 0.0;
  double dgT[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) dgT[c] = (c > lane && c < j) ? Pw[c * (c + 1) / 2 + lane] * dinv : 0.0;
#pragma unroll
  for (int c = 31; c >= 0; --c) {
    if (c >= j) continue;
    const double xc = __shfl_sync(FULL, xv, c);
    xv = fma(-dgT[c], xc, xv);
  }
  if (va) x[lane] = xv;
  __syncwarp();
}
#endif

__global__ void __launch_bounds__(32 * IR_MAXW) BV_CAT(impact_rows_kernel_nc, BV_NC)(const Dims D, const DevPtrs P, const double* __restrict__ gscratch,
                                                                                    const int gsize, const int kz, const int jlo, const int nrows,
                                                                                    const int wstride) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double colbuf[IR_MAXW][64];
  __shared__ int bad[IR_MAXW];
  const bool SV = D.sv != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long job = (long long)blockIdx.x * nw + warp;
  if (job >= (long long)D.U * nrows) return;
  const int u = (int)(job / nrows), j = jlo + (int)(job % nrows);
  const int k = D.k, q = D.q;
  double* Pw = sm + (size_t)warp * wstride;
  double* r = Pw + ((j * (j + 1) / 2 + 1) & ~1);
  double* z = r + j;
  double* x = z + j;
  if (lane == 0) bad[warp] = 0;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const double* G = gscratch + (size_t)u * gsize;
  const int id = j * (j - 1) / 2;
  const double inv_dj = SV ? 1.0 : 1.0 / P.diag[(size_t)u * k + j];
  if (SV) {
    const double* Gj = G + rows_row_off(j);
    for (int p = lane; p < j * (j + 1) / 2; p += 32) Pw[p] = Gj[p];
    for (int b = lane; b < j; b += 32) r[b] = Gj[j * (j + 1) / 2 + b];  // prior_chol_mean = 0 (tri.h:52)
  } else {  // leading j x j block and row j of the full Z'Z (row pitch kz), walked in packed order: the loads of a lane are
            // independent of one another (a row-by-row loop left ONE load in flight per warp: 99 dependent L2 round trips)
    const int np = j * (j + 1) / 2;
    int a = 0, rs = 0;  // row of packed index p and its first packed index
    for (int p0 = lane; p0 < np; p0 += 32 * 8) {
      double v[8];
      int at[8];
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8) {
        const int p = p0 + 32 * q8;
        while (p < np && p >= rs + a + 1) { rs += a + 1; ++a; }
        at[q8] = p < np ? a * kz + (p - rs) : 0;
      }
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8) v[q8] = G[at[q8]];
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8) {
        const int p = p0 + 32 * q8;
        if (p < np) Pw[p] = v[q8] * inv_dj;
      }
    }
    for (int b = lane; b < j; b += 32) r[b] = G[j * kz + b] * inv_dj;
  }
  __syncwarp();
  for (int b = lane; b < j; b += 32) {
    Pw[b * (b + 1) / 2 + b] += P.chol_prec[(size_t)u * q + id + b];
    z[b] = rng.normal_elem(ST_IMPACT_Z, (uint32_t)(id + b));
  }
  __syncwarp();
#if BV_NC == 1
  small_system_solve(Pw, j, r, z, x, colbuf[warp], &bad[warp]);
#else
  chol_blocked_warp<NC>(Pw, j, &bad[warp], colbuf[warp]);
  precision_solve_blocked<NC>(Pw, j, r, z, x);
  __syncwarp();
#endif
  double* Lu = P.L + (size_t)u * k * k;
  for (int b = lane; b < k; b += 32) {
    double val = b == j ? 1.0 : 0.0;
    if (b < j) {
      val = x[b];
      P.contem[(size_t)u * q + id + b] = val;
      const double fit = fabs(val) * P.znorm[(size_t)u * k + b];  // draw_savs, draw.h:398-415
      P.sparse_contem[(size_t)u * q + id + b] = val * fmax(fit - 1.0 / (val * val), 0.0) / fit;
    }
    Lu[j * k + b] = val;
    if (j == 1) Lu[b] = b == 0 ? 1.0 : 0.0;  // row 0 of L (build_inv_lower, draw.h:124-132)
  }
  __syncwarp();
  if (lane == 0 && bad[warp]) atomicOr(&P.flags[u], 2);
}

}  // namespace

// rows jlo .. jhi (jhi <= 32 BV_NC) of every unit
cudaError_t BV_CAT(launch_impact_rows_nc, BV_NC)(const Dims& D, const DevPtrs& P, const double* gscratch, int gsize, int kz, int jlo, int jhi,
                                                 cudaStream_t st) {
  const int nrows = jhi - jlo + 1;
  if (nrows <= 0) return cudaSuccess;
  const int wstride = ((jhi * (jhi + 1) / 2 + 1) & ~1) + 3 * ((jhi + 1) & ~1);
  int nw = IR_MAXW;
  while (nw > 1 && (size_t)nw * wstride * sizeof(double) > 100 * 1024) nw >>= 1;
  const size_t smem = (size_t)nw * wstride * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  const long long jobs = (long long)D.U * nrows;
  cudaFuncSetAttribute(BV_CAT(impact_rows_kernel_nc, BV_NC), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BV_CAT(impact_rows_kernel_nc, BV_NC)<<<(unsigned)((jobs + nw - 1) / nw), nw * 32, smem, st>>>(D, P, gscratch, gsize, kz, jlo, nrows, wstride);
  return cudaGetLastError();
}

}  // namespace bvhar
