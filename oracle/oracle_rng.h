// TEST INFRASTRUCTURE ONLY — CPU oracle for the bvhar Gibbs hot path.  Nothing under oracle/ is
// imported, linked or executed by the product (bvhar_b200/); only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it.
//
// oracle_rng.h — random streams and samplers of the oracle.
//
// Two back ends behind one interface:
//   * PHILOX: the counter-based stream *specification* that the CUDA engine also implements
//     (independently, in bvhar_b200/csrc/rng.cuh).  A draw is addressed by
//     (seed, unit | iteration, stage, element, sub-draw), so the oracle and the device can be
//     compared trajectory-for-trajectory.  Spec: DESIGN.md "Random streams".
//   * MT: std::mt19937 consumed sequentially like the reference's boost::random::mt19937
//     (inst/include/bvhar/src/core/common.h:542-580) with the standard library's distributions —
//     an independent set of sampler algorithms used for the distributional (3 MCSE) checks.
//
// Samplers restated from the reference:
//   inverse Gaussian  — Michael/Schucany/Haas, common.h:99-111
//   GIG               — bounded rejection sampler (Devroye 2014 style), common.h:234-288, 337-376
// Gamma: Marsaglia & Tsang (2000) "A simple method for generating gamma variables" (the reference
// delegates to boost::random::gamma_distribution, an un-vendored dependency: BH >= 1.87,
// DESCRIPTION:44; parity with it is distributional only).
#ifndef BVHAR_ORACLE_RNG_H
#define BVHAR_ORACLE_RNG_H

#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

namespace orc {

// Philox4x32-10 (Salmon et al., SC'11), written out round by round.
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline double u52(uint32_t lo, uint32_t hi) {
  uint64_t x = ((uint64_t)hi << 32) | lo;
  return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);  // (0,1), 52 bits
}

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
  ST_PRIOR_COEF = 16,   // + site 0..7
  ST_PRIOR_CONTEM = 24, // + site 0..7
  ST_MNIW_CHI = 32, ST_MNIW_BART = 33, ST_MNIW_Z = 34, ST_MH_CAND = 35, ST_MH_U = 36  // conjugate Minnesota samplers
};

struct Rng {
  bool philox = true;
  uint32_t key[2] = {0, 0};
  uint32_t iter = 0;
  std::mt19937 mt;

  void seed(uint32_t s, uint32_t unit, bool use_philox) {
    philox = use_philox;
    key[0] = s;
    key[1] = unit;
    mt.seed(s);
  }

  // ---- raw pairs ----
  void uniform_pair(uint32_t stage, uint32_t idx, uint32_t sub, double& a, double& b) {
    if (philox) {
      uint32_t ctr[4] = {idx, sub, stage, iter}, out[4];
      philox4x32_10(ctr, key, out);
      a = u52(out[0], out[1]);
      b = u52(out[2], out[3]);
    } else {
      std::uniform_real_distribution<double> d(0.0, 1.0);
      a = d(mt);
      b = d(mt);
    }
  }
  void normal_pair(uint32_t stage, uint32_t idx, uint32_t sub, double& z0, double& z1) {
    if (philox) {
      double a, b;
      uniform_pair(stage, idx, sub, a, b);
      double r = std::sqrt(-2.0 * std::log(a));
      double th = 6.283185307179586476925286766559 * b;
      z0 = r * std::cos(th);
      z1 = r * std::sin(th);
    } else {
      std::normal_distribution<double> d(0.0, 1.0);
      z0 = d(mt);
      z1 = d(mt);
    }
  }
  // ---- vectors: element e lives in pair e>>1, lane e&1 ----
  void normal_vec(uint32_t stage, uint32_t e0, int len, double* out) {
    if (!philox) {
      std::normal_distribution<double> d(0.0, 1.0);
      for (int i = 0; i < len; ++i) out[i] = d(mt);
      return;
    }
    for (int i = 0; i < len; ++i) {
      uint32_t e = e0 + (uint32_t)i;
      double z0, z1;
      normal_pair(stage, e >> 1, 0, z0, z1);
      out[i] = (e & 1u) ? z1 : z0;
    }
  }
  void uniform_vec(uint32_t stage, uint32_t e0, int len, double* out) {
    if (!philox) {
      std::uniform_real_distribution<double> d(0.0, 1.0);
      for (int i = 0; i < len; ++i) out[i] = d(mt);
      return;
    }
    for (int i = 0; i < len; ++i) {
      uint32_t e = e0 + (uint32_t)i;
      double a, b;
      uniform_pair(stage, e >> 1, 0, a, b);
      out[i] = (e & 1u) ? b : a;
    }
  }
  // ---- scalar site draws ----
  double uniform(uint32_t stage, uint32_t idx, uint32_t sub = 0) {
    double a, b;
    if (philox) { uniform_pair(stage, idx, sub, a, b); return a; }
    std::uniform_real_distribution<double> d(0.0, 1.0);
    return d(mt);
  }
  double normal(uint32_t stage, uint32_t idx, uint32_t sub = 0) {
    double a, b;
    if (philox) { normal_pair(stage, idx, sub, a, b); return a; }
    std::normal_distribution<double> d(0.0, 1.0);
    return d(mt);
  }
  // Gamma(shape, scale), common.h:552-555 semantics.
  double gamma(uint32_t stage, uint32_t idx, double shape, double scale, uint32_t variant = 0) {
    if (!philox) {
      std::gamma_distribution<double> d(shape, scale);
      return d(mt);
    }
    const uint32_t vb = variant << 24;
    double boost = 1.0, a = shape;
    if (a < 1.0) {
      double u = uniform(stage, idx, vb | 0xFFFFFFu);
      boost = std::pow(u, 1.0 / a);
      a += 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / std::sqrt(9.0 * d);
    for (uint32_t t = 0;; ++t) {
      double x = normal(stage, idx, vb | (2u * t));
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      double u = uniform(stage, idx, vb | (2u * t + 1u));
      if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) return d * v * boost * scale;
      if (t > 100000u) return std::numeric_limits<double>::quiet_NaN();
    }
  }
  double beta(uint32_t stage, uint32_t idx, double s1, double s2) {
    // boost::random::beta_distribution is X/(X+Y) with two gammas as well.
    double x = gamma(stage, idx, s1, 1.0, 0);
    double y = gamma(stage, idx, s2, 1.0, 1);
    return x / (x + y);
  }
  double bernoulli(uint32_t stage, uint32_t idx, double p) { return uniform(stage, idx, 0) < p ? 1.0 : 0.0; }
  // discrete_distribution on un-normalised weights (common.h:577-580); NaN weights count as 0.
  int categorical(uint32_t stage, uint32_t idx, const double* w, int n) {
    double tot = 0;
    for (int i = 0; i < n; ++i) if (w[i] == w[i]) tot += w[i];
    double target = uniform(stage, idx, 0) * tot, cum = 0;
    for (int i = 0; i < n; ++i) {
      if (w[i] == w[i]) cum += w[i];
      if (cum > target) return i;
    }
    return n - 1;
  }
  // common.h:99-111 (alpha = mean, beta = shape)
  double invgauss(uint32_t stage, uint32_t idx, double mean, double shape) {
    double z = normal(stage, idx, 0);
    double c = mean / (2.0 * shape);
    double w = mean * z * z;  // alpha * chi2(1)
    // The smaller root, evaluated LITERALLY as common.h:104-106 writes it.  The expression cancels catastrophically when
    // w >> shape (relative error ~ 2^-53 w / shape, it can even come out <= 0), and that rounding noise is load-bearing:
    // with the cancellation-free form mean * 4 shape w / (w + r)^2 the Dirichlet-Laplace contemporaneous block (q = 3)
    // collapses into its absorbing state (global -> DBL_MIN -> NaN within ~400 sweeps), whereas the literal form keeps
    // kicking collapsed chains out again, which is what the reference's posterior reflects.  Do not "fix".
    double cand = mean + c * (w - std::sqrt(w * (4.0 * shape + w)));
    double u = uniform(stage, idx, 1);
    if (u < mean / (mean + cand)) return cand;
    return mean * mean / cand;
  }
  // common.h:234-288: GIG(p, a, b), density ∝ x^(p-1) exp(-(a x + b/x)/2); rand.h:220-223 call order.
  double gig(uint32_t stage, uint32_t idx, double p_, double a_, double b_) {
    const double abs_p = std::fabs(p_);
    const double omega = std::sqrt(a_ * b_);
    const double alpha = std::sqrt(omega * omega + abs_p * abs_p) - abs_p;
    auto psi = [&](double x) { return -alpha * (std::cosh(x) - 1.0) - abs_p * (std::exp(x) - x - 1.0); };
    auto dpsi = [&](double x) { return -alpha * std::sinh(x) - abs_p * (std::exp(x) - 1.0); };
    double t = 1.0, s = 1.0;
    double lc = -psi(1.0);
    if (lc >= 0.5 && lc <= 2.0) t = 1.0;
    else if (lc > 2.0) t = std::sqrt(2.0 / (alpha + abs_p));
    else if (lc < 0.5) t = std::log(4.0 / (alpha + 2.0 * abs_p));
    lc = -psi(-1.0);
    if (lc >= 0.5 && lc <= 2.0) s = 1.0;
    else if (lc > 2.0) s = std::sqrt(4.0 / (alpha * std::cosh(1.0) + abs_p));
    else if (lc < 0.5)
      s = std::fmin(1.0 / abs_p, std::log(1.0 + 1.0 / alpha + std::sqrt(1.0 / (alpha * alpha) + 2.0 / alpha)));
    const double eta = -psi(t), zeta = -dpsi(t), theta = -psi(-s), xi = dpsi(-s);
    const double p = 1.0 / xi, r = 1.0 / zeta;
    const double td = t - r * eta, sd = s - p * theta, q = td + sd;
    double cand = 0;
    for (uint32_t it = 0;; ++it) {
      double u, v, w, dummy;
      if (philox) {
        uniform_pair(stage, idx, 2u * it, u, v);
        uniform_pair(stage, idx, 2u * it + 1u, w, dummy);
      } else {
        u = uniform(stage, idx); v = uniform(stage, idx); w = uniform(stage, idx);
      }
      if (u < q / (p + q + r)) cand = -sd + q * v;
      else if (u < (q + r) / (p + q + r)) cand = td - r * std::log(v);
      else cand = -sd + p * std::log(v);
      double chi;
      if (cand >= -sd && cand <= td) chi = 1.0;
      else if (cand > td) chi = std::exp(-eta - zeta * (cand - t));
      else chi = std::exp(-theta + xi * (cand + s));
      if (!(w * chi > std::exp(psi(cand)))) break;
      if (it > 100000u) return std::numeric_limits<double>::quiet_NaN();
    }
    cand = (abs_p / omega + std::sqrt(1.0 + abs_p * abs_p / (omega * omega))) * std::exp(cand);
    return p_ > 0 ? cand * std::sqrt(b_ / a_) : std::sqrt(b_ / a_) / cand;
  }
};

// draw.h:99-103
inline void cut_param(double& x) {
  if (x < std::numeric_limits<double>::min() || std::isnan(x)) x = std::numeric_limits<double>::min();
}

}  // namespace orc
#endif
