// kernels_setup.cu — one-off device-side setup and the forecast kernel.
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"
#include <math_constants.h>

namespace bvhar {

// XX[t][p(a,b)] = x_ta * x_tb for a >= b: the shared right operand of the G1 contraction
// (every chain of every window reuses it; rolling / expanding windows are row ranges).
__global__ void __launch_bounds__(256) build_xx_kernel(const Dims D, const double* __restrict__ X, double* __restrict__ XX) {
  const int t = blockIdx.x;
  const double* x = X + (size_t)t * D.ldX;
  double* o = XX + (size_t)t * D.ldXX;
  for (int p = threadIdx.x; p < D.ldXX; p += blockDim.x) {
    double v = 0.0;
    if (p < D.Pp) {
      int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
      while (tri_idx(a + 1, 0) <= p) ++a;
      while (tri_idx(a, 0) > p) --a;
      const int b = p - tri_idx(a, 0);
      v = x[a] * x[b];
    }
    o[p] = v;
  }
}
cudaError_t launch_build_xx(const Dims& D, const DevPtrs& P, double* XX, cudaStream_t st) {
  build_xx_kernel<<<D.T, 256, 0, st>>>(D, P.X, XX);
  return cudaGetLastError();
}

// Per-chain log-volatility inits arrive as n x k column-major blocks ([chain][series][t], what `lvol` is in
// param_init, config.h:410-415); the engine keeps them time-major across the window's units.  Tiled transpose.
__global__ void __launch_bounds__(256) scatter_lvol_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int M,
                                                           int ldM, int nsrc) {
  __shared__ double tile[32][33];
  const int m0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, t = t0 + tx;
    tile[r][tx] = (m < M && t < n) ? src[nsrc ? (size_t)m * nsrc + t : (size_t)m] : 0.0;  // nsrc == 0: one value per series
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, m = m0 + tx;
    if (t < n && m < M) dst[(size_t)t * ldM + m] = tile[tx][r];
  }
}
// src: [M = C*k][nsrc] (series-major rows of length nsrc >= n), dst: [n][ldM]
cudaError_t launch_scatter_lvol(const double* src, double* dst, int n, int M, int ldM, int nsrc, cudaStream_t st) {
  scatter_lvol_kernel<<<dim3((M + 31) / 32, (n + 31) / 32), 256, 0, st>>>(src, dst, n, M, ldM, nsrc);
  return cudaGetLastError();
}

// Posterior mean of every column of a record over all units and kept draws, NaN entries skipped (the SAVS records hold
// 0/0 = NaN where a coefficient is exactly zero, draw.h:417-430): on-device counterpart of the colMeans / nanmean that
// the reference front ends run on the returned draws (R/var-bayes.R:488-617).  One block per column chunk; fixed
// summation order (deterministic).
__global__ void __launch_bounds__(256) record_mean_kernel(const double* __restrict__ rec, int U, int cap, int cols, int rows, int thin,
                                                          double* __restrict__ out) {
  __shared__ double ssum[256];
  __shared__ double scnt[256];
  const int col = blockIdx.x;
  double s = 0.0, n = 0.0;
  const long long total = (long long)U * rows;
  for (long long e = threadIdx.x; e < total; e += blockDim.x) {
    const int u = (int)(e / rows), r = (int)(e - (long long)u * rows);
    const double v = rec[((size_t)u * cap + (size_t)r * thin) * cols + col];
    if (v == v) { s += v; n += 1.0; }
  }
  ssum[threadIdx.x] = s; scnt[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[col] = ssum[0] / scnt[0];
}
cudaError_t launch_record_mean(const double* rec, int U, int cap, int cols, int rows, int thin, double* out, cudaStream_t st) {
  record_mean_kernel<<<cols, 256, 0, st>>>(rec, U, cap, cols, rows, thin, out);
  return cudaGetLastError();
}

// Density forecast of one (unit, draw) per thread: McmcForecaster::forecastOut fc.h:162-190,
// RegForecaster fc.h:206-222, SvForecaster fc.h:239-262, computeMean fc.h:300-302 / 338-340.
// Philox key (seed_forecast[unit], unit), iteration = draw index, stages ST_FORECAST_H / _Y,
// element h * k + j.
__global__ void __launch_bounds__(128) forecast_kernel(const ForecastArgs A) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = A.k, nrow = A.nrow, S = A.S, p = A.p, step = A.step;
  const int dd = p * k + (A.mean ? 1 : 0);
  // per-thread workspace in shared memory: last_pvec (dd), tmp ((p-1)k), post_mean (k), sv (k), z (k), hx (har_r)
  const int per = dd + (p - 1) * k + 3 * k + (A.is_vhar ? A.har_r : 0);
  double* ws = sm + (size_t)threadIdx.x * per;
  double* last_pvec = ws;
  double* tmp = last_pvec + dd;
  double* pm = tmp + (p - 1) * k;
  double* svu = pm + k;
  double* zn = svu + k;
  double* hx = zn + k;
  if (i < S) {
    auto at = [&](const RecView& v, int col) { return v.p[(long long)u * v.us + (long long)i * v.sd + (long long)col * v.sc]; };
    const double* last = A.last_obs + (size_t)u * p * k;
    const Philox rng(A.seed[u], A.unit_base + (uint32_t)u, (uint32_t)i);
    for (int l = 0; l < p; ++l)
      for (int j = 0; j < k; ++j) last_pvec[l * k + j] = last[(size_t)j * p + (p - 1 - l)];
    if (A.mean) last_pvec[dd - 1] = 1.0;
    for (int j = 0; j < k; ++j) pm[j] = last_pvec[j];
    for (int e = 0; e < (p - 1) * k; ++e) tmp[e] = last_pvec[k + e];
    for (int j = 0; j < k; ++j) svu[j] = A.sv_model ? at(A.diag, j) : sqrt(at(A.diag, j));
    for (int h = 0; h < step; ++h) {
      for (int e = 0; e < (p - 1) * k; ++e) last_pvec[k + e] = tmp[e];
      for (int j = 0; j < k; ++j) last_pvec[j] = pm[j];
      if (A.is_vhar) {
        for (int r = 0; r < A.har_r; ++r) {
          double s = 0.0;
          for (int c = 0; c < A.har_c; ++c) {
            const double hv = A.har[(size_t)c * A.har_r + r];
            if (hv != 0.0) s += hv * last_pvec[c];
          }
          hx[r] = s;
        }
      }
      const double* xin = A.is_vhar ? hx : last_pvec;
      if (A.act) {
        // Select forecasters (fc.h:372-374, 409-411): mean = x' (activity_graph o coef_mat); the graph is the activity
        // VECTOR reshaped column-major to (nrow + mean) x k (fc.h:356), i.e. graph(r, j) = act[j * (nrow + mean) + r]
        const double* act = A.act + (size_t)u * A.ncoef;
        const int df = nrow + (A.mean ? 1 : 0);
        for (int j = 0; j < k; ++j) {
          double s = 0.0;
          for (int r = 0; r < nrow; ++r) s += act[j * df + r] * at(A.coef, j * nrow + r) * xin[r];
          if (A.mean) s += act[j * df + nrow] * at(A.cst, j) * xin[nrow];
          pm[j] = s;
        }
      } else {
        for (int j = 0; j < k; ++j) {
          double s = 0.0;
          for (int r = 0; r < nrow; ++r) s += at(A.coef, j * nrow + r) * xin[r];
          if (A.mean) s += at(A.cst, j) * xin[nrow];
          pm[j] = s;
        }
      }
      if (A.sv_model) {
        if (A.sv)
          for (int j = 0; j < k; ++j) svu[j] += rng.normal_elem(ST_FORECAST_H, (uint32_t)(h * k + j)) * sqrt(at(A.sigh, j));
        for (int j = 0; j < k; ++j) zn[j] = rng.normal_elem(ST_FORECAST_Y, (uint32_t)(h * k + j)) * exp(svu[j] / 2);
      } else {
        for (int j = 0; j < k; ++j) zn[j] = rng.normal_elem(ST_FORECAST_Y, (uint32_t)(h * k + j)) * svu[j];
      }
      // post_mean += L^-1 e  (unit-lower solve, fc.h:168); L row j = contem[j(j-1)/2 ...]
      for (int j = 0; j < k; ++j) {
        double s = zn[j];
        const int id = j * (j - 1) / 2;
        for (int m = 0; m < j; ++m) s -= at(A.contem, id + m) * zn[m];
        zn[j] = s;
      }
      for (int j = 0; j < k; ++j) {
        pm[j] += zn[j];
        A.out[((size_t)u * S * k + (size_t)i * k + j) * step + h] = pm[j];
      }
      if (A.get_lpl && A.valid) {  // fc.h:221 / 261, signs as written
        double quad = 0.0, lg = 0.0;
        for (int j = 0; j < k; ++j) {
          double s = pm[j] - A.valid[j];
          const int id = j * (j - 1) / 2;
          for (int m = 0; m < j; ++m) s += at(A.contem, id + m) * (pm[m] - A.valid[m]);
          if (A.sv_model) { s *= exp(-svu[j] / 2); lg += svu[j] / 2; }
          else { s /= svu[j]; lg += log(svu[j]); }
          quad += s * s;
        }
        const double val = lg - k * log(2 * 3.14159265358979323846) / 2 - quad / 2;
        if (A.lpl_draw) A.lpl_draw[((size_t)u * step + h) * S + i] = val;  // summed in a fixed order by lpl_reduce_kernel
      }
      for (int e = 0; e < (p - 1) * k; ++e) tmp[e] = last_pvec[e];
    }
  }
}
// lpl[u][h] = mean over the S draws of the per-draw log predictive likelihood (fc.h:92-96), summed in an order that depends
// on S only (strided per-thread partial sums, then a fixed tree): bit-identical on any GPU count and for any block size of
// the forecast kernel (a floating-point atomicAdd here made the last bits of the LPL depend on the scheduling).
__global__ void __launch_bounds__(128) lpl_reduce_kernel(const double* __restrict__ lpl_draw, int S, double* __restrict__ lpl) {
  __shared__ double part[128];
  const double* src = lpl_draw + (size_t)blockIdx.x * S;
  double s = 0.0;
  for (int i = threadIdx.x; i < S; i += 128) s += src[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) lpl[blockIdx.x] = part[0] / S;
}
// RegRecords::computeActivity (config.h:747-758): per record column, the two order statistics that
// boost::accumulators::tail_quantile<left / right> return with the whole sample cached (BH >= 1.87, un-vendored:
// "the largest of the ceil(n a) smallest" / "the smallest of the ceil(n (1 - a)) largest" samples, NaN when that count
// is not below the cache size), then 0 when they have opposite signs, else 1.1.  One CTA per (column, unit): the S
// draws of the column are sorted ascending by a bitonic network in shared memory (S <= ACT_SMEM_MAX), or - for longer
// chains - ranked by counting in a global scratch row.
constexpr int ACT_SMEM_MAX = 16384;
__global__ void __launch_bounds__(256) activity_kernel(const RecView coef, const RecView cst, int N, int S, int n_lo, int n_hi, int ncoef,
                                                       double* __restrict__ act, double* __restrict__ scratch) {
  extern __shared__ __align__(16) double sm[];
  const int col = blockIdx.x, u = blockIdx.y;
  const RecView& v = col < N ? coef : cst;
  const int c = col < N ? col : col - N;
  const double* base = v.p + (long long)u * v.us + (long long)c * v.sc;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double lo = qnan, hi = qnan;
  if (S <= ACT_SMEM_MAX) {
    int P2 = 1;
    while (P2 < S) P2 <<= 1;
    for (int i = threadIdx.x; i < P2; i += blockDim.x) {
      double x = i < S ? base[(long long)i * v.sd] : CUDART_INF;
      sm[i] = x == x ? x : CUDART_INF;  // (NaN draws - a flagged chain - sort last)
    }
    __syncthreads();
    for (int kk = 2; kk <= P2; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < P2; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const double a = sm[i], b = sm[l];
            const bool up = (i & kk) == 0;
            if ((a > b) == up) { sm[i] = b; sm[l] = a; }
          }
        }
        __syncthreads();
      }
    if (n_lo >= 1 && n_lo < S) lo = sm[n_lo - 1];
    if (n_hi >= 1 && n_hi < S) hi = sm[S - n_hi];
  } else {
    double* row = scratch + ((size_t)u * ncoef + col) * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const double x = base[(long long)i * v.sd];
      row[i] = x == x ? x : CUDART_INF;
    }
    __syncthreads();
    __shared__ double pick[2];
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const double x = row[i];
      int rank = 0;  // position of element i in the stable ascending order
      for (int m = 0; m < S; ++m) {
        const double y = row[m];
        rank += (y < x) || (y == x && m < i);
      }
      if (rank == n_lo - 1) pick[0] = x;
      if (rank == S - n_hi) pick[1] = x;
    }
    __syncthreads();
    if (n_lo >= 1 && n_lo < S) lo = pick[0];
    if (n_hi >= 1 && n_hi < S) hi = pick[1];
  }
  if (threadIdx.x == 0) act[(size_t)u * ncoef + col] = (lo * hi < 0.0) ? 0.0 : 1.1;
}
size_t activity_scratch_doubles(int n_units, int ncoef, int S) { return S <= ACT_SMEM_MAX ? 0 : (size_t)n_units * ncoef * S; }
cudaError_t launch_activity(const RecView& coef, const RecView& cst, int n_units, int N, int kc, int S, int n_lo, int n_hi, double* act,
                            double* scratch, cudaStream_t st) {
  int P2 = 1;
  while (P2 < S) P2 <<= 1;
  const size_t smem = S <= ACT_SMEM_MAX ? (size_t)P2 * sizeof(double) : 0;
  cudaFuncSetAttribute(activity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ACT_SMEM_MAX * sizeof(double)));
  activity_kernel<<<dim3(N + kc, n_units), 256, smem, st>>>(coef, cst, N, S, n_lo, n_hi, N + kc, act, scratch);
  return cudaGetLastError();
}

cudaError_t launch_forecast(const ForecastArgs& A, cudaStream_t st) {
  const int dd = A.p * A.k + (A.mean ? 1 : 0);
  const int per = dd + (A.p - 1) * A.k + 3 * A.k + (A.is_vhar ? A.har_r : 0);
  int threads = 128;
  while (threads > 32 && (size_t)threads * per * sizeof(double) > 200 * 1024) threads >>= 1;
  const size_t smem = (size_t)threads * per * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaFuncSetAttribute(forecast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ForecastArgs B = A;
  B.lpl_draw = nullptr;
  const bool want_lpl = A.lpl && A.get_lpl && A.valid;
  if (want_lpl) {
    cudaError_t e = cudaMallocAsync(&B.lpl_draw, (size_t)A.n_units * A.step * A.S * sizeof(double), st);
    if (e != cudaSuccess) return e;
  }
  forecast_kernel<<<dim3((A.S + threads - 1) / threads, A.n_units), threads, smem, st>>>(B);
  if (want_lpl) {
    lpl_reduce_kernel<<<(unsigned)((size_t)A.n_units * A.step), 128, 0, st>>>(B.lpl_draw, A.S, A.lpl);
    cudaFreeAsync(B.lpl_draw, st);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
