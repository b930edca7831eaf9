// gemm_tn.cuh — grouped fp64 "TN" contraction on the tensor pipe (DMMA mma.sync.m8n8k4.f64).
//
//   C_g[M x N] = epilogue( A_g^T B_g ),   A_g stored K_g x M row-major, B_g stored K_g x N row-major
//
// One launch covers every window group g (blockIdx.z).  This single kernel carries the three
// time-dimension contractions of a sweep (DESIGN.md "Kernels"):
//   G1  S  = Dinv' * XX      SV-weighted Gram  X' diag(1/s_i^2) X  for every (unit, equation)
//                            (replaces the Kronecker-design Gram of triangular.h:286 + draw.h:380)
//   G2  Cm = YLD' * X        X' diag(1/s_i^2) (Y L_i')             (triangular.h:297, draw.h:382)
//   G3  Z  = Y - XT' * A     latent_innov = y - x * coef_mat        (triangular.h:342)
// tcgen05.mma has no f64 kind, so the FP64 tensor path on sm_100a is the legacy DMMA mma.sync.
//
// Tiling: 128 x 128 x 16 CTA tile, 8 warps (2 x 4), 64 x 32 warp tile = 8 x 4 DMMA fragments,
// cp.async (LDGSTS) 3-stage pipeline, smem row pitch BM+4 doubles => conflict-free fragment loads.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bvhar {

enum GemmEpilogue { EPI_STORE = 0, EPI_Y_MINUS = 1 };

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;        // K used when Kg == nullptr
  int lda, ldb, ldc;
  int groups;
  const long long* a_off;  // per group element offsets (nullptr => 0)
  const long long* b_off;
  const long long* c_off;
  const int* Kg;           // per group K
  const int* Mg;           // per group M (nullptr => M); grid.y covers max M
  int epi;
  const double* Y;         // EPI_Y_MINUS: C[r][c] = Y[y_off[g] + r * ldy + (c % ymod)] - acc
  const long long* y_off;
  int ldy, ymod;
  int aligned16;           // all of A/B offsets and leading dimensions are even (16-byte chunks legal)
  // split-K (small batches: fewer tiles than SMs).  ksplit > 1: the K range of every tile is cut into ksplit slices that run
  // as separate work items and store their partial products to ws + slice * ws_stride (same layout as C); the caller then
  // sums the slices in a fixed order (launch_splitk_reduce).  EPI_STORE only.
  int ksplit;
  double* ws;
  long long ws_stride;
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 3, GPAD = 4;
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_SMEM = (size_t)GSTAGES * GBK * ((GBM + GPAD) + (GBN + GPAD)) * sizeof(double);

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// The contraction kernel itself lives in ONE translation unit, gemm_tn.cu (it used to be instantiated by every file
// that included this header: ten copies, a 41 MB library and a 4-minute rebuild).
cudaError_t launch_dgemm_tn(const GemmArgs& g, cudaStream_t st);
// K slices for a contraction of `tiles` output tiles and K <= kmax on a device with `sms` SMs (1 = no split)
int gemm_choose_ksplit(long long tiles, int kmax, int sms);
// C[i] = sum over the slices of ws[s * ws_stride + i], i < count, slices added in order
cudaError_t launch_splitk_reduce(const GemmArgs& g, long long count, cudaStream_t st);

}  // namespace bvhar
