// kernels_stable.cu - the stability filter of the forecasters (subsetStable, config.h:884-914, 1003-1034; is_stable /
// build_companion / root_unitcircle, draw.h:1265-1295) for a batch of coefficient draws on the device.
//
// The reference computes the eigenvalues of the VAR(1) companion matrix of every draw and keeps the draw when the
// largest modulus is < threshold.  A nonsymmetric eigensolver is a poor fit for a batch of thousands of 200..440-wide
// matrices on a GPU; the same DECISION is reached with the DMMA contraction of gemm_tn.cuh by repeated squaring of
// B = companion / threshold:
//     rho(B) < 1   <=>   ||B^N||_F < 1 for some N          (=>: Gelfand;  <=: rho^N <= ||B^N||)
//     rho(B) > 1   <=    |trace(B^N)| > m for some N       (trace(B^N) = sum lambda_i^N)
// Every squaring is followed by a Frobenius re-normalisation whose logarithm is accumulated, so nothing over/underflows;
// a draw is decided as soon as one of the two certificates holds.  Draws still undecided after STAB_MAX_SQUARINGS
// squarings (N = 2^40: |1 - rho| below 1e-11, or a pathological rotation) count as unstable.
#include "engine.cuh"
#include "kernels.h"
#include "block_linalg.cuh"

namespace bvhar {

// B = companion(coef) / threshold and its transpose, row-major m x m with pitch ldm; one CTA per draw.
// coef: draw i at coef + i * ld, alpha[j * nrow + r] = A(r, j)  (triangular.h:291-292).  har: optional 3k x (p k)
// column-major with leading dimension ldh (VHAR: VAR coefficients = har' A, draw.h:1290).
__global__ void __launch_bounds__(256) stab_build_kernel(const double* __restrict__ coef, long long ld, int k, int nrow, int p,
                                                         const double* __restrict__ har, int ldh, double inv_thr, double* R, double* Rt,
                                                         int ldm, double* logacc, int* state) {
  const int m = k * p;
  const long long mat = blockIdx.x;
  const double* a = coef + mat * ld;
  double* r = R + (size_t)mat * m * ldm;
  double* rt = Rt + (size_t)mat * m * ldm;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int row = e / m, col = e - row * m;
    double v = 0.0;
    if (row < k) {  // top rows: coef_mat' (build_companion, draw.h:1268)
      if (har) {
        for (int q3 = 0; q3 < nrow; ++q3) v = fma(har[(size_t)col * ldh + q3], a[(size_t)row * nrow + q3], v);
      } else {
        v = a[(size_t)row * nrow + col];
      }
    } else if (col == row - k) {
      v = 1.0;
    }
    v *= inv_thr;
    r[(size_t)row * ldm + col] = v;
    rt[(size_t)col * ldm + row] = v;
  }
  if (threadIdx.x == 0) { logacc[mat] = 0.0; state[mat] = 0; }
}

// After a squaring: Frobenius norm and trace of the new power, the two certificates, re-normalisation of both copies.
// state: 0 undecided, 1 stable, 2 unstable.  kg[mat] = 0 switches the draw's contraction off once it is decided.
__global__ void __launch_bounds__(256) stab_norm_kernel(double* R, double* Rt, int m, int ldm, double* logacc, int* state, int* kg,
                                                        int* undecided) {
  __shared__ double scratch[40];
  const long long mat = blockIdx.x;
  if (state[mat] != 0) return;
  double* r = R + (size_t)mat * m * ldm;
  double* rt = Rt + (size_t)mat * m * ldm;
  double ss = 0.0, tr = 0.0;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int row = e / m, col = e - row * m;
    const double v = r[(size_t)row * ldm + col];
    ss = fma(v, v, ss);
    if (row == col) tr += v;
  }
  ss = block_sum(ss, scratch);
  tr = block_sum(tr, scratch);
  const double la = 2.0 * logacc[mat];           // log of the scale the squared (normalised) matrix stands for
  const double lf = 0.5 * log(ss);                // log ||.||_F of the stored matrix
  int st = 0;
  if (!(ss == ss) || !(ss < 1.0e300)) st = 2;     // NaN / overflow: not a usable draw
  else if (ss == 0.0 || la + lf < 0.0) st = 1;    // ||B^N||_F < 1
  else if (fabs(tr) > 0.0 && la + log(fabs(tr)) > log((double)m)) st = 2;  // |trace(B^N)| > m
  if (st == 0) {
    const double inv = exp(-lf);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int row = e / m, col = e - row * m;
      r[(size_t)row * ldm + col] *= inv;
      rt[(size_t)row * ldm + col] *= inv;
    }
  }
  if (threadIdx.x == 0) {
    logacc[mat] = la + lf;
    state[mat] = st;
    if (st != 0) kg[mat] = 0;
    else atomicAdd(undecided, 1);
  }
}

cudaError_t launch_stab_build(const double* coef, long long ld, long long n_mat, int k, int nrow, int p, const double* har, int ldh,
                              double inv_thr, double* R, double* Rt, int ldm, double* logacc, int* state, cudaStream_t st) {
  stab_build_kernel<<<(unsigned)n_mat, 256, 0, st>>>(coef, ld, k, nrow, p, har, ldh, inv_thr, R, Rt, ldm, logacc, state);
  return cudaGetLastError();
}
cudaError_t launch_stab_norm(double* R, double* Rt, long long n_mat, int m, int ldm, double* logacc, int* state, int* kg, int* undecided,
                             cudaStream_t st) {
  stab_norm_kernel<<<(unsigned)n_mat, 256, 0, st>>>(R, Rt, m, ldm, logacc, state, kg, undecided);
  return cudaGetLastError();
}

}  // namespace bvhar
