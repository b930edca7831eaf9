// gemm_tn.cu - the grouped fp64 "TN" contraction kernel (DMMA) declared in gemm_tn.cuh: G1 / G2 / G3 of the sweep, the
// X'X / X'Y set-up products of the LDLT models and the companion-matrix squarings of the stability filter.
#include "gemm_tn.cuh"

namespace bvhar {

// Loads a GBK x 128 tile (rows k0.., cols c0..) of a row-major matrix into smem [GBK][128 + GPAD].
template <bool ALIGNED16>
__device__ __forceinline__ void gemm_load_tile(double* sm, const double* __restrict__ g, int ld, int k0, int c0, int K,
                                               int ncols, int tid) {
  if (ALIGNED16) {
#pragma unroll
    for (int i = 0; i < (GBK * 64) / GEMM_THREADS; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int row = c >> 6, col = (c & 63) << 1;
      const bool ok = (k0 + row < K) && (c0 + col < ncols);  // ncols rounded up to even by the caller
      const double* src = ok ? g + (size_t)(k0 + row) * ld + c0 + col : g;
      cp_async16(sm + row * (128 + GPAD) + col, src, ok ? 16 : 0);
    }
  } else {
#pragma unroll
    for (int i = 0; i < (GBK * 128) / GEMM_THREADS; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int row = c >> 7, col = c & 127;
      const bool ok = (k0 + row < K) && (c0 + col < ncols);
      const double* src = ok ? g + (size_t)(k0 + row) * ld + c0 + col : g;
      cp_async8(sm + row * (128 + GPAD) + col, src, ok ? 8 : 0);
    }
  }
}

// One k-tile of the warp's 64 x 32 block: m-tiles 2 i + wr (wr = warp row 0 | 1, interleaved so that a partial row
// tile - e.g. rows 128..199 of the M = C k = 200 rows of a 4-chain window at k = 50 - splits its valid m-tiles evenly
// over the two warp rows), n-tiles j.  NMT = m-tiles this warp owns, a compile-time count: a row tile with 72 valid rows
// issues 5/8 of the DMMAs of a full one (predicated-off DMMAs still occupy the tensor pipe, so the count must be static).
template <int NMT>
__device__ __forceinline__ void gemm_ktile(double (&acc)[8][4][2], const double* as, const double* bs, int wr, int wn0, int fr, int fc) {
#pragma unroll
  for (int kk = 0; kk < GBK; kk += 4) {
    double af[NMT], bf[4];
#pragma unroll
    for (int i = 0; i < NMT; ++i) af[i] = as[(kk + fc) * (GBM + GPAD) + (2 * i + wr) * 8 + fr];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * (GBN + GPAD) + wn0 + j * 8 + fr];
#pragma unroll
    for (int i = 0; i < NMT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

// PART = false: every row tile is full (M a multiple of 128, e.g. the 1 024-chain C3 batch): the plain loop, m-tiles
// wr * 8 + i.  PART = true: row tiles may be partial: interleaved m-tiles and a static per-warp m-tile count.
template <bool ALIGNED16, bool PART>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_tn_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  double* As = reinterpret_cast<double*>(gemm_smem_raw);
  double* Bs = As + GSTAGES * GBK * (GBM + GPAD);
  const int grp = blockIdx.z;
  const int K = g.Kg ? g.Kg[grp] : g.K;
  const int M = g.Mg ? g.Mg[grp] : g.M;
  if ((int)blockIdx.y * GBM >= M) return;
  const double* A = g.A + (g.a_off ? g.a_off[grp] : 0);
  const double* B = g.B + (g.b_off ? g.b_off[grp] : 0);
  double* C = g.C + (g.c_off ? g.c_off[grp] : 0);
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 2, wn0 = (warp & 3) * 32;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row (0..7) / k index (0..3)
  const int Mld = ALIGNED16 ? ((M + 1) & ~1) : M;
  const int Nld = ALIGNED16 ? ((g.N + 1) & ~1) : g.N;
  const int mt_valid = m0 + GBM > M ? (M - m0 + 7) / 8 : 16;  // valid m-tiles of this row tile
  const int nmt = PART ? (mt_valid - wr + 1) / 2 : 8;         // ... of which this warp row owns 2 i + wr < mt_valid

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + GBK - 1) / GBK;
#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) {
      gemm_load_tile<ALIGNED16>(As + s * GBK * (GBM + GPAD), A, g.lda, s * GBK, m0, K, Mld, tid);
      gemm_load_tile<ALIGNED16>(Bs + s * GBK * (GBN + GPAD), B, g.ldb, s * GBK, n0, K, Nld, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot freed at iteration kt - 1
      const int nx = kt + GSTAGES - 1;
      if (nx < nk) {
        const int s = nx % GSTAGES;
        gemm_load_tile<ALIGNED16>(As + s * GBK * (GBM + GPAD), A, g.lda, nx * GBK, m0, K, Mld, tid);
        gemm_load_tile<ALIGNED16>(Bs + s * GBK * (GBN + GPAD), B, g.ldb, nx * GBK, n0, K, Nld, tid);
      }
      cp_async_commit();
    }
    const double* as = As + (kt % GSTAGES) * GBK * (GBM + GPAD);
    const double* bs = Bs + (kt % GSTAGES) * GBK * (GBN + GPAD);
    if (!PART) {
#pragma unroll
      for (int kk = 0; kk < GBK; kk += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) af[i] = as[(kk + fc) * (GBM + GPAD) + wr * 64 + i * 8 + fr];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * (GBN + GPAD) + wn0 + j * 8 + fr];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
      switch (nmt) {  // warp-uniform
        case 8: gemm_ktile<8>(acc, as, bs, wr, wn0, fr, fc); break;
        case 7: gemm_ktile<7>(acc, as, bs, wr, wn0, fr, fc); break;
        case 6: gemm_ktile<6>(acc, as, bs, wr, wn0, fr, fc); break;
        case 5: gemm_ktile<5>(acc, as, bs, wr, wn0, fr, fc); break;
        case 4: gemm_ktile<4>(acc, as, bs, wr, wn0, fr, fc); break;
        case 3: gemm_ktile<3>(acc, as, bs, wr, wn0, fr, fc); break;
        case 2: gemm_ktile<2>(acc, as, bs, wr, wn0, fr, fc); break;
        case 1: gemm_ktile<1>(acc, as, bs, wr, wn0, fr, fc); break;
        default: break;
      }
    }
  }
  cp_async_wait<0>();

  const long long yoff = (g.epi == EPI_Y_MINUS && g.y_off) ? g.y_off[grp] : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (PART ? (2 * i + wr) * 8 : wr * 64 + i * 8) + fr;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wn0 + j * 8 + fc * 2;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (g.epi == EPI_Y_MINUS) {
        if (col < g.N) v0 = g.Y[yoff + (size_t)row * g.ldy + (col % g.ymod)] - v0;
        if (col + 1 < g.N) v1 = g.Y[yoff + (size_t)row * g.ldy + ((col + 1) % g.ymod)] - v1;
      }
      double* dst = C + (size_t)row * g.ldc + col;
      if (col + 1 < g.N && ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      } else {
        if (col < g.N) dst[0] = v0;
        if (col + 1 < g.N) dst[1] = v1;
      }
    }
  }
}

cudaError_t launch_dgemm_tn(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, g.groups);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dgemm_tn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    configured = true;
  }
  const bool part = g.Mg != nullptr || (g.M % GBM) != 0;
  if (g.aligned16) {
    if (part) dgemm_tn_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
    else dgemm_tn_kernel<true, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
  } else {
    if (part) dgemm_tn_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
    else dgemm_tn_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
