// kernels_mniw.cu — the conjugate Minnesota samplers: McmcMniw::doPosteriorDraws
// (bayes/mniw/minnesota.h:241-278) and MhMinnesota::doPosteriorDraws (minnesota.h:413-510), i.e. sim_mn_iw
// (math/random.h:177-211) once per iteration, plus the random-walk Metropolis step on (lambda, psi).
//
// The reference runs `num_iter` sequential iterations per chain on one mt19937.  Nothing in an iteration depends on the
// previous one: the posterior (mean, precision, IW scale, shape) is fixed, and the Metropolis proposal is centred on
// `prevprior`, which the reference never moves (minnesota.h:447 - replicated on purpose).  With the addressed Philox
// stream (rng.cuh: key = (seed_chain, chain), counter = (element, sub-draw, stage, iteration)) every (chain, kept
// iteration) is therefore an independent work item:
//
//   mniw_draw_kernel   one CTA per kept draw (persistent loop over draws).  Bartlett factor of the inverse Wishart
//                      (chi-square diagonal by Marsaglia-Tsang, normal off-diagonal), Sigma = C C' with
//                      C = chol(scale) Q'^-1, its Cholesky factor, then A = M + R^-1 (Z U_v) with R'R = prec factored
//                      once per call and shared by every draw (L2-resident).  The k x k work matrices live in shared
//                      memory; the d x k normal matrix is overwritten in place by Z U_v and then by the solution of the
//                      d-step back substitution (all columns at once, several threads per column).
//   mh_accept_kernel   one thread per (chain, iteration): proposal, log posterior ratio, accept flag.
//   mh_state_kernel    one thread per (chain, kept iteration): (lambda, psi) = the proposal of the last accepted
//                      iteration at or before it (backward scan over the accept flags), or the initial value.
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"

namespace bvhar {

namespace {

constexpr int MN_THREADS = 256;

__global__ void __launch_bounds__(MN_THREADS) mniw_draw_kernel(const MniwArgs A) {
  extern __shared__ __align__(16) double sm[];
  const int k = A.k, d = A.d, tid = threadIdx.x;
  double* Bm = sm;            // [k][k] Bartlett factor B = Q' (lower), later Sigma and its lower Cholesky factor
  double* Ym = Bm + k * k;    // [k][k] B^-1, then C = chol(scale) B^-1 in place
  double* Wsm = Ym + k * k;   // [k][d] (column c of the d x k matrix at c * d) when it fits in shared memory
  __shared__ int bad;
  int tpc = 1;  // threads per column in the back substitution (power of two, tpc * k <= MN_THREADS, <= 32)
  while (tpc < 32 && 2 * tpc * k <= MN_THREADS) tpc <<= 1;

  const long long total = (long long)A.chains * A.rows;
  for (long long job = blockIdx.x; job < total; job += gridDim.x) {
    const int chain = (int)(job / A.rows), r = (int)(job % A.rows);
    const uint32_t t = (uint32_t)(A.num_burn + 1 + r * A.thin);  // mcmc_step of this kept draw (minnesota.h:246)
    const Philox rng(A.seeds[chain], A.ubase + (uint32_t)chain, t);
    double* W = A.wscratch ? A.wscratch + (size_t)blockIdx.x * d * k : Wsm;
    if (tid == 0) bad = 0;
    __syncthreads();  // the previous draw is done with the shared arrays

    // 1. sim_iw_tri (random.h:177-200): upper Bartlett factor Q, stored transposed (B = Q', lower)
    for (int e = tid; e < k * k; e += MN_THREADS) {
      const int i = e / k, j = e - i * k;
      double v = 0.0;
      if (j == i) v = sqrt(rng.gamma(ST_MNIW_CHI, (uint32_t)i, 0.5 * (A.shape - (double)i), 2.0));  // chi^2(shape - i)
      else if (j < i) v = rng.normal_elem(ST_MNIW_BART, (uint32_t)(j * k + i));                      // Q(j, i), j < i
      Bm[e] = v;
    }
    // the d x k standard normals of sim_mn (random.h:163-168), element (i, j) at index i * k + j
    for (int p = tid; p < (d * k + 1) / 2; p += MN_THREADS) {
      double z0, z1;
      rng.normal_pair(ST_MNIW_Z, (uint32_t)p, 0, z0, z1);
      const int e0 = 2 * p, i0 = e0 / k, j0 = e0 - i0 * k;
      W[j0 * d + i0] = z0;
      if (e0 + 1 < d * k) {
        const int i1 = (e0 + 1) / k, j1 = e0 + 1 - i1 * k;
        W[j1 * d + i1] = z1;
      }
    }
    __syncthreads();
    // 2. Y = B^-1 (lower), one thread per column
    if (tid < k) {
      const int c = tid;
      for (int i = 0; i < k; ++i) {
        double v = 0.0;
        if (i == c) v = 1.0 / Bm[i * k + i];
        else if (i > c) {
          double s = 0.0;
          for (int m = c; m < i; ++m) s = fma(Bm[i * k + m], Ym[m * k + c], s);
          v = -s / Bm[i * k + i];
        }
        Ym[i * k + c] = v;
      }
    }
    __syncthreads();
    // 3. C = chol(scale) * B^-1 (lower) in place of B^-1: thread c owns column c and walks the rows downwards, so every
    //    Y[m][c], m <= i, it reads is still the original
    if (tid < k) {
      const int c = tid;
      for (int i = k - 1; i >= c; --i) {
        double s = 0.0;
        for (int m = c; m <= i; ++m) s = fma(A.Ls[i * k + m], Ym[m * k + c], s);
        Ym[i * k + c] = s;
      }
    }
    __syncthreads();
    // 4. Sigma = C C' (random.h:206) -> Bm and the sigma record
    for (int e = tid; e < k * k; e += MN_THREADS) {
      const int i = e / k, j = e - i * k;
      const int mx = i < j ? i : j;
      double s = 0.0;
      for (int m = 0; m <= mx; ++m) s = fma(Ym[i * k + m], Ym[j * k + m], s);
      Bm[e] = s;
      A.sigma[((size_t)chain * k * k + (size_t)j * k + i) * A.rows + r] = s;
    }
    __syncthreads();
    // 5. chol_scale_v = Sigma.llt().matrixU() (random.h:161): lower factor L_v in place, U_v = L_v'
    for (int c = 0; c < k; ++c) {
      const double p = Bm[c * k + c];
      __syncthreads();
      const double inv = rsqrt(p);
      if (tid == 0) {
        Bm[c * k + c] = p * inv;
        if (!(p > 0.0)) bad = 1;
      }
      for (int a = c + 1 + tid; a < k; a += MN_THREADS) Bm[a * k + c] *= inv;
      __syncthreads();
      const int rem = k - c - 1;
      for (int e = tid; e < rem * rem; e += MN_THREADS) {
        const int a = c + 1 + e / rem, b = c + 1 + e % rem;
        if (b <= a) Bm[a * k + b] = fma(-Bm[a * k + c], Bm[b * k + c], Bm[a * k + b]);
      }
      __syncthreads();
    }
    // 6. W <- Z U_v in place, columns in descending order (column c needs Z columns j <= c only)
    for (int i = tid; i < d; i += MN_THREADS) {
      for (int c = k - 1; c >= 0; --c) {
        double s = 0.0;
        for (int j = 0; j <= c; ++j) s = fma(W[j * d + i], Bm[c * k + j], s);
        W[c * d + i] = s;
      }
    }
    __syncthreads();
    // 7. X = U_u^-1 W with U_u = prec.llt().matrixU() = R (random.h:170): back substitution, all columns at once
    {
      const int c = tid / tpc, sub = tid - c * tpc;
      const bool on = c < k;
      for (int i = d - 1; i >= 0; --i) {
        double s = 0.0;
        if (on) {
          const double* Ri = A.Rfull + (size_t)i * d;
          const double* Wc = W + (size_t)c * d;
          for (int m = i + 1 + sub; m < d; m += tpc) s = fma(Ri[m], Wc[m], s);
        }
        for (int o = tpc >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (on && sub == 0) W[(size_t)c * d + i] = (W[(size_t)c * d + i] - s) / A.Rfull[(size_t)i * d + i];
        __syncthreads();
      }
    }
    // 8. coef = mean + X, vectorised column-major (minnesota.h:128)
    for (int e = tid; e < d * k; e += MN_THREADS)
      A.alpha[((size_t)chain * d * k + e) * A.rows + r] = A.mean[e] + W[e];
    if (tid == 0 && bad) atomicOr(A.flags, 1);
  }
}

// log of the Gamma(shape, scale) density, Rf_dgamma(x, shape, scale, TRUE) (common.h:464-466); -inf outside the support
__device__ __forceinline__ double mh_gamma_logdens(double x, double shp, double scl) {
  if (!(x > 0.0)) return -INFINITY;
  return (shp - 1.0) * log(x) - x / scl - lgamma(shp) - shp * log(scl);
}
// invgamma_dens(x, shp, scl, true), evaluated literally as common.h:502-517 writes it (density first, then its log)
__device__ __forceinline__ double mh_invgamma_logdens(double x, double shp, double scl) {
  return log(pow(scl, shp) * pow(x, -shp - 1.0) * exp(-scl / x) / tgamma(shp));
}
// candprior = prevprior + z' chol(gaussian_variance).U (minnesota.h:447, random.h:9-21); G row-major upper
__device__ __forceinline__ double mh_candidate(const Philox& rng, const double* prev, const double* G, int np, int j) {
  double v = prev[j];
  for (int i = 0; i <= j; ++i) v = fma(rng.normal_elem(ST_MH_CAND, (uint32_t)i), G[i * np + j], v);
  return v;
}

__global__ void mh_accept_kernel(const MhArgs M) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)M.chains * M.num_iter) return;
  const int chain = (int)(id / M.num_iter);
  const uint32_t t = (uint32_t)(id % M.num_iter) + 1u;
  const int np = 1 + M.k;
  const double* prev = M.init_par + (size_t)chain * np;
  const double* G = M.chol_var + (size_t)chain * np * np;
  const Philox rng(M.seeds[chain], M.ubase + (uint32_t)chain, t);
  // jointdens_hyperparam (draw.h:59-72) at the candidate minus at prevprior: compute_logml and the multivariate-gamma
  // constants are called with identical arguments on both sides (minnesota.h:448-455) and cancel
  const double lam = mh_candidate(rng, prev, G, np, 0);
  double num = mh_gamma_logdens(lam, M.gamma_shp, 1.0 / M.gamma_rate);
  double den = mh_gamma_logdens(prev[0], M.gamma_shp, 1.0 / M.gamma_rate);
  bool neg = false;
  for (int j = 1; j < np; ++j) {
    const double psi = mh_candidate(rng, prev, G, np, j);
    if (psi < 0.0) neg = true;  // invgamma_dens STOPs: "'x' should be larger than 0." (common.h:503-505)
    num += mh_invgamma_logdens(psi, M.invgam_shp, M.invgam_scl);
    den += mh_invgamma_logdens(prev[j], M.invgam_shp, M.invgam_scl);
  }
  if (neg) atomicOr(M.flags, 2);
  const double u = rng.uniform(ST_MH_U, 0, 0);
  M.accept_all[id] = (log(u) < fmin(num - den, 0.0)) ? 1 : 0;  // minnesota.h:456
}

__global__ void mh_state_kernel(const MhArgs M) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)M.chains * M.rows) return;
  const int chain = (int)(id / M.rows), r = (int)(id % M.rows);
  const int np = 1 + M.k;
  const int t = M.num_burn + 1 + r * M.thin;
  const unsigned char* acc = M.accept_all + (size_t)chain * M.num_iter;
  int s = t;
  while (s >= 1 && !acc[s - 1]) --s;  // last accepted iteration <= t
  const double* prev = M.init_par + (size_t)chain * np;
  M.accept[(size_t)chain * M.rows + r] = acc[t - 1];
  if (s < 1) {
    M.lambda[(size_t)chain * M.rows + r] = prev[0];
    for (int j = 1; j < np; ++j) M.psi[((size_t)chain * M.k + (j - 1)) * M.rows + r] = prev[j];
    return;
  }
  const double* G = M.chol_var + (size_t)chain * np * np;
  const Philox rng(M.seeds[chain], M.ubase + (uint32_t)chain, (uint32_t)s);
  M.lambda[(size_t)chain * M.rows + r] = mh_candidate(rng, prev, G, np, 0);
  for (int j = 1; j < np; ++j) M.psi[((size_t)chain * M.k + (j - 1)) * M.rows + r] = mh_candidate(rng, prev, G, np, j);
}

}  // namespace

size_t mniw_smem(int k, int d, bool w_in_smem) { return ((size_t)2 * k * k + (w_in_smem ? (size_t)d * k : 0)) * sizeof(double); }

cudaError_t launch_mniw(const MniwArgs& A, int grid, size_t smem, cudaStream_t st) {
  if (A.k > MN_THREADS) return cudaErrorInvalidConfiguration;  // one thread per column in steps 2, 3 and 7
  cudaFuncSetAttribute(mniw_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mniw_draw_kernel<<<grid, MN_THREADS, smem, st>>>(A);
  return cudaGetLastError();
}

cudaError_t launch_mh(const MhArgs& M, cudaStream_t st) {
  const long long n1 = (long long)M.chains * M.num_iter, n2 = (long long)M.chains * M.rows;
  mh_accept_kernel<<<(unsigned)((n1 + 127) / 128), 128, 0, st>>>(M);
  mh_state_kernel<<<(unsigned)((n2 + 127) / 128), 128, 0, st>>>(M);
  return cudaGetLastError();
}

}  // namespace bvhar
