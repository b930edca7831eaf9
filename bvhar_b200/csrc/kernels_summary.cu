// kernels_summary.cu — posterior summaries of a record on the device: what the reference's
// front ends compute on the returned draws - colMeans, the PIP colMeans(alpha_sparse != 0) (R/var-bayes.R:503-531) - and
// what their users ask the `posterior` package for next (summarise_draws: sd, quantiles, rhat, ess), without moving the
// draws to the host.  One CTA per record column; the column's draws [chain][draw] are staged in shared memory when they
// fit, otherwise every pass re-reads them from the record buffer.
//
//   mean, sd        over all chains and kept draws (NaN entries - SAVS 0/0 - skipped; sd with n - 1)
//   q_lo, median, q_hi   numpy / R type-7 quantiles at level / 2, 1 / 2, 1 - level / 2: exact order statistics by an
//                   8-bit radix select on the order-preserving integer image of the doubles (8 passes, any sample size)
//   rhat            split-R-hat (Gelman et al., BDA3; posterior::rhat_basic): chains split in halves
//   ess             posterior::ess_basic / Stan: autocovariances of the split chains, Geyer's initial positive and
//                   monotone sequences, tau truncated at 1 / log10(N)
//   pip             share of non-zero entries (NaN skipped)
#include "engine.cuh"
#include "kernels.h"

namespace bvhar {

namespace {

constexpr int SUM_THREADS = 256;
constexpr int RHO_MAX = 2048;  // autocorrelation lags kept by the ESS estimator (a chain that needs more has ESS ~ 0 anyway)
constexpr int SUM_NQ = 6;  // order statistics selected concurrently (lower / upper neighbour of three quantiles)

__device__ __forceinline__ unsigned long long ord_key(double x) {  // monotone map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ double block_reduce_sum(double v, double* red) {
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < SUM_THREADS / 32; ++w) s += red[w];  // fixed order: deterministic
  return s;
}

__global__ void __launch_bounds__(SUM_THREADS) record_summary_kernel(const double* __restrict__ rec, int U, int cap, int cols, int S, int thin,
                                                                    double level, int in_smem, double* __restrict__ out) {
  extern __shared__ __align__(16) double xs[];  // [U][S] when in_smem, then the 2 U means of the split chains
  __shared__ double red[SUM_THREADS / 32];
  __shared__ unsigned hist[SUM_NQ][256];
  __shared__ unsigned long long prefix[SUM_NQ];
  __shared__ long long want[SUM_NQ];
  __shared__ double sh[8];
  __shared__ double rho[RHO_MAX];
  const int col = blockIdx.x, tid = threadIdx.x;
  const long long n = (long long)U * S;
  double* cmean = xs + (in_smem ? n : 0);
  auto X = [&](long long e) -> double {
    if (in_smem) return xs[e];
    const long long u = e / S, r = e - u * S;
    return rec[((size_t)u * cap + (size_t)r * thin) * cols + col];
  };
  if (in_smem) {
    for (long long e = tid; e < n; e += SUM_THREADS) {
      const long long u = e / S, r = e - u * S;
      xs[e] = rec[((size_t)u * cap + (size_t)r * thin) * cols + col];
    }
    __syncthreads();
  }
  // ---- moments, pip (NaN skipped)
  double s = 0.0, cnt = 0.0, nz = 0.0;
  for (long long e = tid; e < n; e += SUM_THREADS) {
    const double v = X(e);
    if (v == v) { s += v; cnt += 1.0; nz += v != 0.0 ? 1.0 : 0.0; }
  }
  s = block_reduce_sum(s, red); cnt = block_reduce_sum(cnt, red); nz = block_reduce_sum(nz, red);
  const double mean = s / cnt;
  double ss = 0.0;
  for (long long e = tid; e < n; e += SUM_THREADS) {
    const double v = X(e);
    if (v == v) ss += (v - mean) * (v - mean);
  }
  ss = block_reduce_sum(ss, red);
  const bool complete = cnt == (double)n;
  double* o = out + (size_t)col * 8;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (tid == 0) { o[0] = mean; o[1] = cnt > 1.0 ? sqrt(ss / (cnt - 1.0)) : qnan; o[7] = nz / cnt; }
  if (!complete || n < 1) {  // order statistics and chain diagnostics are undefined with missing entries
    if (tid == 0) { o[2] = o[3] = o[4] = o[5] = o[6] = qnan; }
    return;
  }
  // ---- quantiles: radix select of the 0-based ranks floor(h), floor(h) + 1 with h = (n - 1) p
  const double probs[3] = {level / 2, 0.5, 1.0 - level / 2};
  if (tid < SUM_NQ) {
    const double h = (double)(n - 1) * probs[tid >> 1];
    long long lo = (long long)floor(h);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    want[tid] = (tid & 1) ? (lo + 1 > n - 1 ? n - 1 : lo + 1) : lo;
    prefix[tid] = 0ULL;
  }
  __syncthreads();
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int e = tid; e < SUM_NQ * 256; e += SUM_THREADS) (&hist[0][0])[e] = 0u;
    __syncthreads();
    const unsigned long long himask = shift == 56 ? 0ULL : (~0ULL << (shift + 8));
    for (long long e = tid; e < n; e += SUM_THREADS) {
      const unsigned long long key = ord_key(X(e));
#pragma unroll
      for (int q = 0; q < SUM_NQ; ++q)
        if ((key & himask) == prefix[q]) atomicAdd(&hist[q][(key >> shift) & 255ULL], 1u);
    }
    __syncthreads();
    if (tid < SUM_NQ) {
      long long r = want[tid];
      int b = 0;
      for (; b < 255; ++b) {
        const long long c = hist[tid][b];
        if (r < c) break;
        r -= c;
      }
      want[tid] = r;
      prefix[tid] |= (unsigned long long)b << shift;
    }
    __syncthreads();
  }
  if (tid < 3) {
    const double h = (double)(n - 1) * probs[tid];
    const double fl = floor(h);
    const double lo = ord_val(prefix[2 * tid]), hi = ord_val(prefix[2 * tid + 1]);
    o[2 + tid] = lo + (h - fl) * (hi - lo);
  }
  // ---- split chains: m = 2 U half chains of length n2 = floor(S / 2); the second half starts at ceil(S / 2)
  const int n2 = S / 2, off2 = (S + 1) / 2, m = 2 * U;
  if (n2 < 2) {
    if (tid == 0) { o[5] = qnan; o[6] = qnan; }
    return;
  }
  auto XC = [&](int c, int i) -> double { return X((long long)(c >> 1) * S + ((c & 1) ? off2 : 0) + i); };
  // chain means / variances: W, B.  One warp per split chain (warp reductions only), one block reduction per quantity.
  const int lane = tid & 31, warp = tid >> 5;
  double l_mean = 0.0, l_mean2 = 0.0, l_var = 0.0;
  for (int c = warp; c < m; c += SUM_THREADS / 32) {
    double a = 0.0;
    for (int i = lane; i < n2; i += 32) a += XC(c, i);
    for (int of = 16; of > 0; of >>= 1) a += __shfl_xor_sync(0xffffffffu, a, of);
    const double cm = a / n2;
    double v = 0.0;
    for (int i = lane; i < n2; i += 32) { const double dd = XC(c, i) - cm; v += dd * dd; }
    for (int of = 16; of > 0; of >>= 1) v += __shfl_xor_sync(0xffffffffu, v, of);
    if (lane == 0) { cmean[c] = cm; l_mean += cm; l_mean2 += cm * cm; l_var += v / (n2 - 1); }
  }
  const double sum_mean = block_reduce_sum(l_mean, red), sum_mean2 = block_reduce_sum(l_mean2, red), sum_var = block_reduce_sum(l_var, red);
  const double W = sum_var / m;
  const double gm = sum_mean / m;
  const double var_means = m > 1 ? (sum_mean2 - m * gm * gm) / (m - 1) : 0.0;
  const double var_plus = W * (n2 - 1) / n2 + var_means;  // (n - 1) / n W + B / n with B = n var(means)
  if (tid == 0) o[5] = sqrt(var_plus / W);
  // ---- ESS: rho_t = 1 - (W - mean_c acov_c(t)) / var_plus, acov biased (divided by n2)
  auto acov_mean = [&](int t) -> double {  // mean over the split chains of the lag-t autocovariance
    double tot = 0.0;
    for (int c = warp; c < m; c += SUM_THREADS / 32) {
      const double cm = cmean[c];
      double v = 0.0;
      for (int i = lane; i + t < n2; i += 32) v += (XC(c, i) - cm) * (XC(c, i + t) - cm);
      for (int of = 16; of > 0; of >>= 1) v += __shfl_xor_sync(0xffffffffu, v, of);
      if (lane == 0) tot += v / n2;
    }
    return block_reduce_sum(tot, red) / m;
  };
  // the estimator follows posterior:::.ess (= Stan's), literally: mean_var = W, var_plus as above, Geyer's initial positive
  // sequence over pairs of lags, the "improved estimate" term rho[max_t], then the initial monotone sequence.  Every
  // thread runs the same control flow (the reductions return the same value to all); thread 0 keeps the rho array.
  const double mean_var = W;
  if (tid == 0) { for (int i = 0; i < RHO_MAX; ++i) rho[i] = 0.0; }
  double rho_even = 1.0, rho_odd = 1.0 - (mean_var - acov_mean(1)) / var_plus;
  if (tid == 0) { rho[0] = rho_even; rho[1] = rho_odd; }
  int t = 0;
  while (t < n2 - 5 && (rho_even + rho_odd) == (rho_even + rho_odd) && (rho_even + rho_odd) > 0.0 && t + 4 < RHO_MAX) {
    t += 2;
    rho_even = 1.0 - (mean_var - acov_mean(t)) / var_plus;
    rho_odd = 1.0 - (mean_var - acov_mean(t + 1)) / var_plus;
    if (tid == 0 && (rho_even + rho_odd) >= 0.0) { rho[t] = rho_even; rho[t + 1] = rho_odd; }
  }
  const int max_t = t;
  if (tid == 0) {
    if (rho_even > 0.0) rho[max_t] = rho_even;
    int tt = 0;
    while (tt <= max_t - 4) {
      tt += 2;
      if (rho[tt] + rho[tt + 1] > rho[tt - 2] + rho[tt - 1]) {
        rho[tt] = (rho[tt - 2] + rho[tt - 1]) / 2;
        rho[tt + 1] = rho[tt];
      }
    }
    double tau = -1.0 + rho[max_t];
    const int upto = max_t > 1 ? max_t : 1;
    for (int i = 0; i < upto; ++i) tau += 2.0 * rho[i];
    const double N = (double)m * n2;
    tau = fmax(tau, 1.0 / log10(N));
    o[6] = N / tau;
  }
  (void)sh;
}

}  // namespace

cudaError_t launch_record_summary(const double* rec, int U, int cap, int cols, int S, int thin, double level, double* out, cudaStream_t st) {
  const size_t need = (size_t)U * S * sizeof(double), means = (size_t)2 * U * sizeof(double);
  const int in_smem = need + means <= 180 * 1024 ? 1 : 0;
  const size_t smem = (in_smem ? need : 0) + means;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;  // > 12 800 chains
  if (smem > 48 * 1024) cudaFuncSetAttribute(record_summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  record_summary_kernel<<<cols, SUM_THREADS, smem, st>>>(rec, U, cap, cols, S, thin, level, in_smem, out);
  return cudaGetLastError();
}

}  // namespace bvhar
