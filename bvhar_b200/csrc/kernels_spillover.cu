// kernels_spillover.cu - spillover densities from stored draws (McmcSpillover::computeSpillover, spillover.h:48-66;
// convert_var_to_vma / convert_vhar_to_vma / compute_vma_fevd / compute_sp_index / compute_to / _from / _tot / _net,
// math/structural.h:8-146): per draw, Sigma = L^-1 D L^-T, the VMA coefficients W_0..W_{step-1} of the VAR (or of the VAR
// form har' Phi of a VHAR), the normalised generalised FEVD at horizon `step` and the connectedness table built on it.
// One CTA per draw (grid-stride); the k x k work matrices of a CTA live in a global scratch slab (L2-resident).
#include "block_linalg.cuh"
#include "engine.cuh"
#include "kernels.h"

namespace bvhar {

__global__ void __launch_bounds__(256) spillover_kernel(const SpilloverArgs A) {
  __shared__ double scratch[40];
  const int k = A.k, p = A.lag, step = A.step, nrow = A.nrow;
  const int kk = k * k, tid = threadIdx.x, nt = blockDim.x;
  double* slab = A.scratch + (size_t)blockIdx.x * A.slab;
  double* Lm = slab;                      // [k][k] unit lower (build_inv_lower, draw.h:124-132), row-major
  double* sq = Lm + kk;                   // [k][k] sqrt_sig = L^-1 D^(1/2)
  double* cov = sq + kk;                  // [k][k]
  double* numer = cov + kk;               // [k][k]
  double* innov = numer + kk;             // [k] diagonal of sum_i W_i' Sigma W_i
  double* svv = innov + k;                // [k] D^(1/2)
  double* mp = svv + k;                   // [k][k] ma_prod = W_i' Sigma
  double* Av = mp + kk;                   // [p k][k] VAR coefficients (row block l = lag l + 1)
  double* W = Av + (size_t)p * kk;        // [step][k][k] VMA blocks
  for (long long i = blockIdx.x; i < A.n_draws; i += gridDim.x) {
    // draw i = (unit, s): units are e.g. the (window, chain) pairs of device-resident records, or the time points of a
    // dynamic SV spillover (the same coefficient rows, log-volatilities advancing by dim per unit)
    const long long un = i / A.per_unit, sd = i - un * A.per_unit;
    const double* alpha = A.coef + un * A.us_coef + sd * A.ld_coef;
    const double* a = A.contem + un * A.us_contem + sd * A.ld_contem;
    const double* dv = A.sv + un * A.us_sv + sd * A.ld_sv;
    __syncthreads();
    // D^(1/2): updateDiag (config.h:876-881 LDLT, 980-996 SV)
    for (int c = tid; c < k; c += nt) {
      double v;
      if (A.sv_kind == 0) v = sqrt(dv[c]);
      else if (A.sv_kind == 1) v = exp(0.5 * dv[c]);
      else {
        v = 0.0;
        for (int t = 0; t < A.n_time; ++t) v += exp(0.5 * dv[(size_t)t * k + c]);
        v /= A.n_time;
      }
      svv[c] = v;
      innov[c] = 0.0;
    }
    for (int e = tid; e < kk; e += nt) {
      const int r = e / k, c = e - r * k;
      Lm[e] = r == c ? 1.0 : (c < r ? a[r * (r - 1) / 2 + c] : 0.0);
      numer[e] = 0.0;
    }
    // VAR coefficient blocks: A[r][j] = alpha[j nrow + r], VHAR: har' Phi (structural.h:53)
    for (int e = tid; e < p * kk; e += nt) {
      const int r = e / k, j = e - r * k;
      double v = 0.0;
      if (A.har) {
        for (int q3 = 0; q3 < nrow; ++q3) v = fma(A.har[(size_t)r * A.ldh + q3], alpha[(size_t)j * nrow + q3], v);
      } else {
        v = alpha[(size_t)j * nrow + r];
      }
      Av[e] = v;
    }
    __syncthreads();
    // sqrt_sig = L^-1 diag(sv): column c solves L x = sv_c e_c (unit lower)
    for (int c = tid; c < k; c += nt) {
      for (int r = 0; r < c; ++r) sq[r * k + c] = 0.0;
      sq[c * k + c] = svv[c];
      for (int r = c + 1; r < k; ++r) {
        double s = 0.0;
        for (int m2 = c; m2 < r; ++m2) s = fma(Lm[r * k + m2], sq[m2 * k + c], s);
        sq[r * k + c] = -s;
      }
    }
    __syncthreads();
    for (int e = tid; e < kk; e += nt) {
      const int r = e / k, c = e - r * k;
      double s = 0.0;
      for (int m2 = 0; m2 < k; ++m2) s = fma(sq[r * k + m2], sq[c * k + m2], s);
      cov[e] = s;
      W[e] = r == c ? 1.0 : 0.0;  // W_0 = I
    }
    __syncthreads();
    for (int h = 0; h < step; ++h) {
      double* Wh = W + (size_t)h * kk;
      if (h > 0) {  // W_h = sum_{l < min(h, p)} A_l W_{h-l-1}   (structural.h:25-33)
        for (int e = tid; e < kk; e += nt) {
          const int r = e / k, c = e - r * k;
          double s = 0.0;
          const int lmax = min(h, p);
          for (int l = 0; l < lmax; ++l) {
            const double* Al = Av + (size_t)l * kk + (size_t)r * k;
            const double* Wp = W + (size_t)(h - l - 1) * kk + c;
            for (int m2 = 0; m2 < k; ++m2) s = fma(Al[m2], Wp[(size_t)m2 * k], s);
          }
          Wh[e] = s;
        }
        __syncthreads();
      }
      // ma_prod = W_h' Sigma; innov_rr += (ma_prod W_h)_rr; numer += (ma_prod / sqrt(sigma_cc))^2   (structural.h:113-117)
      for (int e = tid; e < kk; e += nt) {
        const int r = e / k, c = e - r * k;
        double s = 0.0;
        for (int m2 = 0; m2 < k; ++m2) s = fma(Wh[(size_t)m2 * k + r], cov[(size_t)m2 * k + c], s);
        mp[e] = s;
        const double v = s / sqrt(cov[(size_t)c * k + c]);
        numer[e] += v * v;
      }
      __syncthreads();
      for (int r = tid; r < k; r += nt) {
        double s = 0.0;
        for (int c = 0; c < k; ++c) s = fma(mp[(size_t)r * k + c], Wh[(size_t)c * k + r], s);
        innov[r] += s;
      }
      __syncthreads();
    }
    // last FEVD block, rows normalised, x 100 (structural.h:118-126); reuse mp for the table
    for (int r = tid; r < k; r += nt) {
      double rs = 0.0;
      for (int c = 0; c < k; ++c) rs += numer[(size_t)r * k + c] / innov[r];
      for (int c = 0; c < k; ++c) mp[(size_t)r * k + c] = 100.0 * (numer[(size_t)r * k + c] / innov[r]) / rs;
    }
    __syncthreads();
    double off = 0.0;
    for (int e = tid; e < kk; e += nt) {
      const int r = e / k, c = e - r * k;
      const double v = mp[e];
      A.connect[((size_t)i * k + c) * k + r] = v;                                        // k x (S k) column-major
      A.net_pairwise[((size_t)i * k + c) * k + r] = (mp[(size_t)c * k + r] - v) / k;     // (S' - S) / k
      if (r != c) off += v;
    }
    for (int c = tid; c < k; c += nt) {
      double to = 0.0, from = 0.0;
      for (int r = 0; r < k; ++r)
        if (r != c) { to += mp[(size_t)r * k + c]; from += mp[(size_t)c * k + r]; }
      A.to[(size_t)i * k + c] = to;      // column sums without the diagonal
      A.from[(size_t)i * k + c] = from;  // row sums without the diagonal
    }
    off = block_sum(off, scratch);
    if (tid == 0) A.tot[i] = off / k;
  }
}

size_t spillover_slab_doubles(int k, int lag, int step) {
  const size_t kk = (size_t)k * k;
  return 5 * kk + 2 * (size_t)k + (size_t)lag * kk + (size_t)step * kk;
}
cudaError_t launch_spillover(const SpilloverArgs& A, int n_cta, cudaStream_t st) {
  spillover_kernel<<<n_cta, 256, 0, st>>>(A);
  return cudaGetLastError();
}

}  // namespace bvhar
