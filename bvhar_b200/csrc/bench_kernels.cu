// bench_kernels.cu — device-resident micro-benchmarks behind the C-ABI, used by bench.py for the
// roofline denominators that MEASURED_PEAKS.json does not carry (fp64 FMA / DMMA issue peaks) and to
// time the contraction kernel alone.
#include <string>
#include <vector>

#include "../../include/bvhar_cuda.h"
#include "gemm_tn.cuh"

using namespace bvhar;

namespace {
// 8 independent accumulator chains per thread, 2 flops per DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 independent m8n8k4 accumulator tiles per warp, 512 flops per DMMA.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma_m8n8k4(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Dependent-issue latency of the fp64 operations the serial chains of the sampler are made of (pivot chains of the warp
// Cholesky, tridiagonal recurrences, substitution steps): one warp, one chain, clock64 around `iters` dependent operations.
// kind 0: DFMA, 1: DMMA m8n8k4 (accumulator chain), 2: rsqrt(double), 3: 1.0 / x, 4: a DFMA chain through a 64-bit shuffle
// (the step of the register-block substitutions).  out[0] = cycles per operation.
// One warp, ONE dependent chain, straight-line: 64 operations per loop trip so that the loop branch does not count.
template <int KIND>
__global__ void __launch_bounds__(32) fp64_latency_kernel(double* out, int trips, double seed) {
  double a = seed, b = 1.0 + 1e-9 * seed, c0 = 0.0, c1 = 0.0;
  const double e = 1e-9 * seed;
  const int src = (threadIdx.x + 1) & 31;
  const long long t0 = clock64();
  for (int it = 0; it < trips; ++it) {
#pragma unroll
    for (int r = 0; r < 64; ++r) {
      if (KIND == 0) a = fma(a, b, e);
      else if (KIND == 1) dmma_m8n8k4(c0, c1, b, a);
      else if (KIND == 2) a = rsqrt(a) + 1.5;
      else if (KIND == 3) a = 1.0 / a + 0.5;
      else if (KIND == 4) a = fma(__shfl_sync(0xffffffffu, a, src), b, e);
      else if (KIND == 5) a = a + e;
      else a = a * b;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / (64.0 * trips); out[1] = a + c0 + c1; }
}
__global__ void fill_kernel(double* p, size_t n, double scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = scale * (double)((i * 2654435761u) % 1000) / 1000.0 + 0.1;
}
}  // namespace

extern "C" {

// mode 0: DFMA, mode 1: DMMA m8n8k4.  Returns achieved TFLOP/s with every SM busy.
int bvhar_cuda_bench_fp64_peak(int device, int mode, double* tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return BVHAR_ERR_CUDA;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return BVHAR_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters);
    else dmma_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flops = mode == 0 ? (double)blocks * threads * iters * 64.0 * 2.0 : (double)blocks * (threads / 32) * iters * 64.0 * 512.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return cudaGetLastError() == cudaSuccess ? 0 : BVHAR_ERR_CUDA;
}

// cycles per dependent fp64 operation (see fp64_latency_kernel); kind 0..6
int bvhar_cuda_bench_fp64_latency(int device, int kind, double* cycles) {
  if (cudaSetDevice(device) != cudaSuccess) return BVHAR_ERR_CUDA;
  double* out;
  if (cudaMalloc(&out, 2 * sizeof(double)) != cudaSuccess) return BVHAR_ERR_CUDA;
  double h[2] = {0.0, 0.0};
  for (int rep = 0; rep < 2; ++rep) {
    switch (kind) {
      case 0: fp64_latency_kernel<0><<<1, 32>>>(out, 400, 1.25); break;
      case 1: fp64_latency_kernel<1><<<1, 32>>>(out, 400, 1.25); break;
      case 2: fp64_latency_kernel<2><<<1, 32>>>(out, 400, 1.25); break;
      case 3: fp64_latency_kernel<3><<<1, 32>>>(out, 400, 1.25); break;
      case 4: fp64_latency_kernel<4><<<1, 32>>>(out, 400, 1.25); break;
      case 5: fp64_latency_kernel<5><<<1, 32>>>(out, 400, 1.25); break;
      default: fp64_latency_kernel<6><<<1, 32>>>(out, 400, 1.25); break;
    }
    if (cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(out); return BVHAR_ERR_CUDA; }
  }
  *cycles = h[0];
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? 0 : BVHAR_ERR_CUDA;
}

// Times the TN contraction kernel on device-resident synthetic operands (CUDA events, `iters` launches).
int bvhar_cuda_bench_dgemm_tn(int device, int64_t M, int64_t N, int64_t K, int iters, double* ms_per_launch) {
  if (cudaSetDevice(device) != cudaSuccess) return BVHAR_ERR_CUDA;
  const int64_t lda = (M + 1) & ~1LL, ldb = (N + 1) & ~1LL, ldc = ldb;
  double *A, *B, *Cc;
  if (cudaMalloc(&A, sizeof(double) * K * lda) != cudaSuccess || cudaMalloc(&B, sizeof(double) * K * ldb) != cudaSuccess ||
      cudaMalloc(&Cc, sizeof(double) * M * ldc) != cudaSuccess)
    return BVHAR_ERR_CUDA;
  fill_kernel<<<1024, 256>>>(A, (size_t)K * lda, 1.0);
  fill_kernel<<<1024, 256>>>(B, (size_t)K * ldb, 0.5);
  GemmArgs g{};
  g.A = A; g.B = B; g.C = Cc; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.lda = (int)lda; g.ldb = (int)ldb; g.ldc = (int)ldc;
  g.groups = 1; g.epi = EPI_STORE; g.aligned16 = 1;
  launch_dgemm_tn(g, 0);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) launch_dgemm_tn(g, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_launch = ms / iters;
  cudaFree(A); cudaFree(B); cudaFree(Cc); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return cudaGetLastError() == cudaSuccess ? 0 : BVHAR_ERR_CUDA;
}
}
