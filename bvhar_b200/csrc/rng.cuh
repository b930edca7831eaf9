// rng.cuh — counter-based random streams and device samplers of the Gibbs engine.
//
// Every draw is addressed, not sequenced: Philox4x32-10 with
//     key     = (seed_chain[window, chain], unit index)
//     counter = (element, sub-draw, stage, iteration)
// so a unit's draws are bit-identical on any GPU, at any GPU count and under any thread
// mapping (BASELINE.json north_star; replaces one boost::random::mt19937 per chain,
// inst/include/bvhar/src/bayes/triangular/triangular.h:207).  Stream spec: DESIGN.md
// "Random streams".  Distributions replaced (core/common.h:542-580, math/random.h:220-230):
// normal, uniform, gamma, beta, Bernoulli, categorical, inverse Gaussian (common.h:99-111) and
// generalized inverse Gaussian (common.h:234-288).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bvhar {

enum Stage : uint32_t {
  ST_COEF_Z = 1,
  ST_IMPACT_Z = 2,
  ST_SV_MIX_U = 3,
  ST_SV_H_Z = 4,
  ST_SV_SIGH = 5,
  ST_SV_H0_Z = 6,
  ST_LDLT_DIAG = 7,
  ST_FORECAST_H = 8,
  ST_FORECAST_Y = 9,
  ST_PRIOR_COEF = 16,
  ST_PRIOR_CONTEM = 24,
  // conjugate Minnesota samplers (kernels_mniw.cu)
  ST_MNIW_CHI = 32,   // Bartlett diagonal: chi-square(shape - i), element i
  ST_MNIW_BART = 33,  // Bartlett off-diagonal normals, element j * k + i (j < i)
  ST_MNIW_Z = 34,     // d x k matrix-normal innovations, element i * k + j
  ST_MH_CAND = 35,    // Metropolis proposal normals, element 0 .. k
  ST_MH_U = 36        // Metropolis acceptance uniform
};

struct Philox {
  uint32_t k0, k1, iter;
  __device__ __forceinline__ Philox(uint32_t seed, uint32_t unit, uint32_t it) : k0(seed), k1(unit), iter(it) {}

  __device__ __forceinline__ uint4 raw(uint32_t stage, uint32_t idx, uint32_t sub) const {
    uint32_t c0 = idx, c1 = sub, c2 = stage, c3 = iter, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  static __device__ __forceinline__ double u52(uint32_t lo, uint32_t hi) {
    const unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);
  }
  __device__ __forceinline__ void uniform_pair(uint32_t stage, uint32_t idx, uint32_t sub, double& a, double& b) const {
    const uint4 r = raw(stage, idx, sub);
    a = u52(r.x, r.y);
    b = u52(r.z, r.w);
  }
  __device__ __forceinline__ void normal_pair(uint32_t stage, uint32_t idx, uint32_t sub, double& z0, double& z1) const {
    double a, b;
    uniform_pair(stage, idx, sub, a, b);
    const double r = sqrt(-2.0 * log(a));
    double s, c;
    sincos(6.283185307179586476925286766559 * b, &s, &c);
    z0 = r * c;
    z1 = r * s;
  }
  // element e of a vector draw: pair e >> 1, lane e & 1
  __device__ __forceinline__ double normal_elem(uint32_t stage, uint32_t e) const {
    double z0, z1;
    normal_pair(stage, e >> 1, 0, z0, z1);
    return (e & 1u) ? z1 : z0;
  }
  __device__ __forceinline__ double uniform_elem(uint32_t stage, uint32_t e) const {
    double a, b;
    uniform_pair(stage, e >> 1, 0, a, b);
    return (e & 1u) ? b : a;
  }
  __device__ __forceinline__ double uniform(uint32_t stage, uint32_t idx, uint32_t sub = 0) const {
    double a, b;
    uniform_pair(stage, idx, sub, a, b);
    return a;
  }
  __device__ __forceinline__ double normal(uint32_t stage, uint32_t idx, uint32_t sub = 0) const {
    double a, b;
    normal_pair(stage, idx, sub, a, b);
    return a;
  }
  // Gamma(shape, scale): Marsaglia-Tsang, attempt t uses sub-draws 2t (normal) and 2t+1 (uniform).
  __device__ double gamma(uint32_t stage, uint32_t idx, double shape, double scale, uint32_t variant = 0) const {
    const uint32_t vb = variant << 24;
    double boost = 1.0, a = shape;
    if (a < 1.0) {
      boost = pow(uniform(stage, idx, vb | 0xFFFFFFu), 1.0 / a);
      a += 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t <= 100000u; ++t) {
      const double x = normal(stage, idx, vb | (2u * t));
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = uniform(stage, idx, vb | (2u * t + 1u));
      if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v * boost * scale;
    }
    return __longlong_as_double(0x7ff8000000000000LL);
  }
  __device__ double beta(uint32_t stage, uint32_t idx, double s1, double s2) const {
    const double x = gamma(stage, idx, s1, 1.0, 0);
    const double y = gamma(stage, idx, s2, 1.0, 1);
    return x / (x + y);
  }
  __device__ __forceinline__ double bernoulli(uint32_t stage, uint32_t idx, double p) const {
    return uniform(stage, idx, 0) < p ? 1.0 : 0.0;
  }
  // Michael-Schucany-Haas, common.h:99-111 (alpha = mean, beta = shape)
  __device__ double invgauss(uint32_t stage, uint32_t idx, double mean, double shape) const {
    const double z = normal(stage, idx, 0);
    const double c = mean / (2.0 * shape);
    const double w = mean * z * z;
    // literal evaluation of common.h:104-106 on purpose: the catastrophic cancellation for w >> shape is part of the
    // reference's behaviour (it keeps the Dirichlet-Laplace chains out of their absorbing state), see oracle_rng.h
    const double cand = mean + c * (w - sqrt(w * (4.0 * shape + w)));
    const double u = uniform(stage, idx, 1);
    if (u < mean / (mean + cand)) return cand;
    return mean * mean / cand;
  }
  // GIG(p, a, b) ∝ x^(p-1) exp(-(a x + b / x) / 2), common.h:234-288, 337-376
  __device__ double gig(uint32_t stage, uint32_t idx, double p_, double a_, double b_) const {
    const double abs_p = fabs(p_);
    const double omega = sqrt(a_ * b_);
    const double alpha = sqrt(omega * omega + abs_p * abs_p) - abs_p;
#define BV_PSI(x) (-alpha * (cosh(x) - 1.0) - abs_p * (exp(x) - (x) - 1.0))
#define BV_DPSI(x) (-alpha * sinh(x) - abs_p * (exp(x) - 1.0))
    double t = 1.0, s = 1.0;
    double lc = -BV_PSI(1.0);
    if (lc >= 0.5 && lc <= 2.0) t = 1.0;
    else if (lc > 2.0) t = sqrt(2.0 / (alpha + abs_p));
    else if (lc < 0.5) t = log(4.0 / (alpha + 2.0 * abs_p));
    lc = -BV_PSI(-1.0);
    if (lc >= 0.5 && lc <= 2.0) s = 1.0;
    else if (lc > 2.0) s = sqrt(4.0 / (alpha * cosh(1.0) + abs_p));
    else if (lc < 0.5) s = fmin(1.0 / abs_p, log(1.0 + 1.0 / alpha + sqrt(1.0 / (alpha * alpha) + 2.0 / alpha)));
    const double eta = -BV_PSI(t), zeta = -BV_DPSI(t), theta = -BV_PSI(-s), xi = BV_DPSI(-s);
    const double p = 1.0 / xi, r = 1.0 / zeta;
    const double td = t - r * eta, sd = s - p * theta, q = td + sd;
    double cand = 0.0;
    bool ok = false;
    for (uint32_t it = 0; it <= 100000u; ++it) {
      double u, v, w, dummy;
      uniform_pair(stage, idx, 2u * it, u, v);
      uniform_pair(stage, idx, 2u * it + 1u, w, dummy);
      if (u < q / (p + q + r)) cand = -sd + q * v;
      else if (u < (q + r) / (p + q + r)) cand = td - r * log(v);
      else cand = -sd + p * log(v);
      double chi;
      if (cand >= -sd && cand <= td) chi = 1.0;
      else if (cand > td) chi = exp(-eta - zeta * (cand - t));
      else chi = exp(-theta + xi * (cand + s));
      if (!(w * chi > exp(BV_PSI(cand)))) { ok = true; break; }
    }
#undef BV_PSI
#undef BV_DPSI
    if (!ok) return __longlong_as_double(0x7ff8000000000000LL);
    cand = (abs_p / omega + sqrt(1.0 + abs_p * abs_p / (omega * omega))) * exp(cand);
    return p_ > 0 ? cand * sqrt(b_ / a_) : sqrt(b_ / a_) / cand;
  }
};

// draw.h:99-103
__device__ __forceinline__ double cut_param(double x) {
  return (x < 2.2250738585072014e-308 || isnan(x)) ? 2.2250738585072014e-308 : x;
}

}  // namespace bvhar
