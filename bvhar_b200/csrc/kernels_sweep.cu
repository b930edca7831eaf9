// kernels_sweep.cu — the per-iteration updaters of the corrected-triangular Gibbs sweep
// (McmcTriangular::doPosteriorDraws, inst/include/bvhar/src/bayes/triangular/triangular.h:102-115)
// as sm_100a kernels.  One kernel (or one contraction + one kernel) per updater:
//
//   stage 3  updateSv      sv_prep_kernel      tri.h:409           elementwise, HBM-bound
//   stage 4  updateCoef    G1/G2 contractions + coef_kernel        tri.h:281-316, draw.h:372-383, 417-430
//   stage 6  updateLatent  G3 contraction                          tri.h:342
//   stage 7  updateImpact  impact_kernel       tri.h:322-336, draw.h:398-415   (+ stage 8 updateChol, draw.h:124-132)
//   stage 9  updateState   SV:   sv_mix_kernel + sv_solve_kernel + sv_finish_kernel   tri.h:400-408, draw.h:440-537
//                          LDLT: ldlt_state_kernel (kernels_impact.cu)                tri.h:371, draw.h:1244-1259
//   records                record_kernel       config.h:850-857, 955-965, 796-804
//
// Arithmetic follows the "efficient" formulation proved equal to the reference's in
// oracle/bvhar_oracle.cpp (update_coef_efficient, varsv_ht_efficient) and DESIGN.md.
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "block_linalg.cuh"
#include <algorithm>
#include <cstdlib>

namespace bvhar {

// ------------------------------------------------------------------ stage 3: updateSv (SV)
// Dinv = exp(-h) = 1 / sqrt_sv^2 and YLD = (Y L')_{t,i} * Dinv, both in [t][m] contraction layout.
// One thread per series m keeps its row of L in registers and walks a chunk of time points; the
// response rows of the chunk are staged in shared memory.
constexpr int SVP_TT = 16;  // time points per block
template <int KMAX>
__global__ void __launch_bounds__(256, (KMAX <= 32 ? 2 : 1)) sv_prep_kernel(const Dims D, const DevPtrs P) {
  extern __shared__ __align__(16) double sm[];
  const int w = blockIdx.z;
  const int n = P.win_n[w], row0 = P.win_row0[w];
  const int t0 = blockIdx.y * SVP_TT;
  if (t0 >= n) return;
  const int k = D.k;
  const int tn = min(SVP_TT, n - t0);
  for (int e = threadIdx.x; e < tn * k; e += blockDim.x) sm[e] = P.Y[(size_t)(row0 + t0) * k + e];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= D.M) return;
  const int c = m / k, i = m - c * k;
  const int u = w * D.C + c;
  double lr[KMAX];
  const double* Lr = P.L + ((size_t)u * k + i) * k;
#pragma unroll
  for (int m2 = 0; m2 < KMAX; ++m2) lr[m2] = (m2 <= i) ? Lr[m2] : 0.0;
  const long long base = P.win_off[w] + (long long)t0 * D.ldM + m;
#pragma unroll 4
  for (int tt = 0; tt < tn; ++tt) {
    const long long at = base + (long long)tt * D.ldM;
    const double dinv = exp(-P.H[at]);
    const double* y = sm + tt * k;
    double yl = 0.0;
#pragma unroll
    for (int m2 = 0; m2 < KMAX; ++m2)
      if (m2 < k) yl = fma(lr[m2], y[m2], yl);
    P.Dinv[at] = dinv;
    P.YLD[at] = yl * dinv;
  }
}

// ------------------------------------------------------------------ stage 4: updateCoef
// One CTA per unit; equations j = 0..k-1 in order (the draw of equation j conditions on this
// sweep's draws of equations < j, tri.h:283, 297):
//   P_j = diag(prec_j) + sum_{i>=j} L_ij^2 S_i
//   b_j = prec_j o mean_j + sum_{i>=j} L_ij (C_i - S_i v_i),   v_i = sum_{m<=i, m != j} L_im a_m
//   a_j = P_j^-1 b_j + chol(P_j)^-T z,  z ~ N(0, I_d)  (Philox stage ST_COEF_Z, element j*d + a)
// SV:   S_i, C_i come from the G1/G2 contractions.  LDLT: S_i = X'X / d_i, C_i = X'Y L_i' / d_i.
// The packed S_i are streamed through shared memory in row chunks of <= CH doubles; the symmetric
// product S_i v is split into a row pass (warp per row) and a column pass (thread per column).
template <bool SV>
__global__ void __launch_bounds__(256) coef_kernel(const Dims D, const DevPtrs P, const int CH, double* vscratch, double* pscratch) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = D.k, d = D.d, Pp = D.Pp, ldM = D.ldM;
  // the packed precision / factor lives in shared memory when it fits, else in a per-unit global (L2-resident) scratch
  double* Pk = pscratch ? pscratch + (size_t)u * ((Pp + 1) & ~1) : sm;
  double* Sbuf = pscratch ? sm : sm + ((Pp + 1) & ~1);
  double* Lsm = Sbuf + CH;
  double* rhs = Lsm + k * k;
  double* outv = rhs + d;
  double* lo = outv + d;
  double* up = lo + d;
  double* zz = up + d;
  double* vs = zz + d;     // scaled vector of the current contribution
  double* wl = vs + d;     // per-i weights (k)
  double* V = vscratch ? vscratch + (size_t)u * 2 * k * d : wl + k;
  double* Asm = V + k * d;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);

  for (int e = tid; e < k * k; e += blockDim.x) Lsm[e] = P.L[(size_t)u * k * k + e];
  double* Ac = P.Acoef + (size_t)w * d * ldM + (size_t)c * k;
  for (int e = tid; e < d * k; e += blockDim.x) {
    const int a = e / k, j = e % k;
    Asm[j * d + a] = Ac[(size_t)a * ldM + j];
  }
  __syncthreads();
  for (int e = tid; e < k * d; e += blockDim.x) {
    const int i = e / d, a = e % d;
    double s = 0.0;
    for (int m = 0; m <= i; ++m) s += Lsm[i * k + m] * Asm[m * d + a];
    V[e] = s;
  }
  __syncthreads();

  const double* prec_u = P.prec + (size_t)u * k * d;
  const double* xn = P.xnorm + (size_t)w * d;
  for (int j = 0; j < k; ++j) {
    // drop equation j from v_i (coef_mat.col(j).setZero(), tri.h:283)
    for (int e = tid; e < (k - j) * d; e += blockDim.x) {
      const int i = j + e / d, a = e % d;
      V[i * d + a] -= Lsm[i * k + j] * Asm[j * d + a];
    }
    for (int p = tid; p < Pp; p += blockDim.x) Pk[p] = 0.0;
    for (int a = tid; a < d; a += blockDim.x) {
      lo[a] = 0.0;
      up[a] = 0.0;
      zz[a] = rng.normal_elem(ST_COEF_Z, (uint32_t)(j * d + a));
    }
    if (!SV) {
      for (int i = j + tid; i < k; i += blockDim.x) wl[i] = Lsm[i * k + j] / P.diag[(size_t)u * k + i];
    }
    __syncthreads();
    const int ni = SV ? (k - j) : 1;
    for (int ii = 0; ii < ni; ++ii) {
      const int i = j + ii;
      const double* Sg;
      double w2;
      __syncthreads();  // passes of the previous contribution are done reading vs
      if (SV) {
        const double lij = Lsm[i * k + j];
        Sg = P.S + ((size_t)u * k + i) * D.ldS;
        w2 = lij * lij;
        for (int a = tid; a < d; a += blockDim.x) vs[a] = lij * V[i * d + a];
      } else {
        Sg = P.XtX + (size_t)w * D.ldS;
        double ws = 0.0;
        for (int i2 = j; i2 < k; ++i2) ws += Lsm[i2 * k + j] * wl[i2];
        w2 = ws;
        for (int a = tid; a < d; a += blockDim.x) {
          double s = 0.0;
          for (int i2 = j; i2 < k; ++i2) s += wl[i2] * V[i2 * d + a];
          vs[a] = s;
        }
      }
      int a0 = 0;
      while (a0 < d) {
        int a1 = a0 + 1;  // rows [a0, a1) with packed length <= CH
        while (a1 < d && tri_idx(a1 + 1, 0) - tri_idx(a0, 0) <= CH) ++a1;
        const int p0 = tri_idx(a0, 0), p1 = tri_idx(a1, 0);
        __syncthreads();  // previous chunk consumed; vs visible
        for (int p = p0 + tid; p < p1; p += blockDim.x) {
          const double s = Sg[p];
          Sbuf[p - p0] = s;
          Pk[p] += w2 * s;
        }
        __syncthreads();
        for (int a = a0 + warp; a < a1; a += (blockDim.x >> 5)) {  // row pass: sum_{b<=a} S_ab v_b
          const double* row = Sbuf + (tri_idx(a, 0) - p0);
          double s = 0.0;
          for (int b = lane; b <= a; b += 32) s += row[b] * vs[b];
          s = warp_sum(s);
          if (lane == 0) lo[a] += s;
        }
        for (int b = tid; b < a1 - 1; b += blockDim.x) {  // column pass: sum_{a>b, a in chunk} S_ab v_a
          double s = 0.0;
          for (int a = max(a0, b + 1); a < a1; ++a) s += Sbuf[tri_idx(a, b) - p0] * vs[a];
          up[b] += s;
        }
        a0 = a1;
      }
    }
    __syncthreads();
    // right-hand side and prior diagonal
    for (int a = tid; a < d; a += blockDim.x) {
      double ct = 0.0;
      if (SV) {
        for (int i = j; i < k; ++i) ct += Lsm[i * k + j] * P.Cm[((size_t)u * k + i) * D.ldX + a];
      } else {
        const double* xty = P.XtY + ((size_t)w * d + a) * k;
        for (int i = j; i < k; ++i) {
          double s = 0.0;
          for (int m = 0; m <= i; ++m) s += xty[m] * Lsm[i * k + m];
          ct += wl[i] * s;
        }
      }
      const double pr = prec_u[j * d + a];
      rhs[a] = pr * P.prior_mean[j * d + a] + ct - lo[a] - up[a];
      Pk[tri_idx(a, a)] += pr;
    }
    chol_packed_block(Pk, d, &bad);
    if (warp == 0) precision_solve_warp(Pk, d, rhs, zz, outv);
    __syncthreads();
    for (int a = tid; a < d; a += blockDim.x) {
      const double val = outv[a];
      Asm[j * d + a] = val;
      Ac[(size_t)a * ldM + j] = val;
      double pen = 0.0;
      if (a < D.nrow) {
        P.alpha[(size_t)u * D.N + j * D.nrow + a] = val;
        pen = P.own_flag[j * D.nrow + a] ? 0.0 : 1.0;
      } else {
        P.cvec[(size_t)u * k + j] = val;
      }
      const double fit = fabs(val) * xn[a];  // draw_mn_savs, draw.h:417-430
      P.sparse_coef[(size_t)u * d * k + j * d + a] = val * fmax(fit - pen / (val * val), 0.0) / fit;
    }
    for (int e = tid; e < (k - j) * d; e += blockDim.x) {
      const int i = j + e / d, a = e % d;
      V[i * d + a] += Lsm[i * k + j] * outv[a];
    }
    __syncthreads();
  }
  if (tid == 0 && bad) atomicOr(&P.flags[u], 1);
}

// ------------------------------------------------------------------ stage 9 (SV): updateState
__constant__ double c_ksc_p[7] = {0.0073, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.2575};       // draw.h:445
__constant__ double c_ksc_m[7] = {-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819}; // draw.h:447
__constant__ double c_ksc_v[7] = {5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261};      // draw.h:450

// Mixture indicators (draw.h:458-468) for every (t, series): y* = log((Z L')^2 + 1e-4) (tri.h:401-402),
// s_t by inverse CDF on the 7-component weights p_c N(y* - h; m_c - 1.2704, v_c); leaves the tridiagonal
// system's diagonal term 1/v_s in Dinv and right-hand-side term (y* - m_s)/v_s in YLD.
// Same thread mapping as sv_prep (row of L in registers); the unit's residual rows are staged in smem.
__constant__ double c_ksc_mm[7] = {-10.12999 - 1.2704, -3.97281 - 1.2704, -8.56686 - 1.2704, 2.77786 - 1.2704,
                                   0.61942 - 1.2704, 1.79518 - 1.2704, -1.08819 - 1.2704};
constexpr float MIX_MARGIN = 2e-4f;
template <int KMAX>
__global__ void __launch_bounds__(256, (KMAX <= 32 ? 2 : 1)) sv_mix_kernel(const Dims D, const DevPtrs P) {
  extern __shared__ __align__(16) double sm[];  // [SVP_TT][blockDim.x span of Z columns]
  __shared__ double s_h2v[7], s_lw[7], s_iv[7];
  __shared__ float s_lwf[7];
  const int w = blockIdx.z;
  const int n = P.win_n[w];
  const int t0 = blockIdx.y * SVP_TT;
  if (t0 >= n) return;
  const int k = D.k;
  if (threadIdx.x < 7) {
    const double v = c_ksc_v[threadIdx.x];
    s_h2v[threadIdx.x] = 0.5 / v;
    s_iv[threadIdx.x] = 1.0 / v;
    s_lw[threadIdx.x] = c_ksc_p[threadIdx.x] / (sqrt(v) * sqrt(2 * 3.14159265358979323846));
    s_lwf[threadIdx.x] = (float)s_lw[threadIdx.x];
  }
  const int tn = min(SVP_TT, n - t0);
  const int mbase = blockIdx.x * blockDim.x;
  const int c0 = mbase / k;                                  // first unit touched by this block
  const int zlo = c0 * k;                                    // first Z column staged
  const int mhi = min(D.M, mbase + (int)blockDim.x);
  const int zhi = min(D.M, ((mhi - 1) / k + 1) * k);         // one past the last Z column staged
  const int zw = zhi - zlo;
  const long long wbase = P.win_off[w] + (long long)t0 * D.ldM;
  for (int e = threadIdx.x; e < tn * zw; e += blockDim.x) {
    const int tt = e / zw, col = e - tt * zw;
    sm[e] = P.Z[wbase + (long long)tt * D.ldM + zlo + col];
  }
  __syncthreads();
  const int m = mbase + threadIdx.x;
  if (m >= D.M) return;
  const int c = m / k, i = m - c * k;
  const int u = w * D.C + c;
  double lr[KMAX];
  const double* Lr = P.L + ((size_t)u * k + i) * k;
#pragma unroll
  for (int m2 = 0; m2 < KMAX; ++m2) lr[m2] = (m2 <= i) ? Lr[m2] : 0.0;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const uint32_t nh = (uint32_t)((n + 1) / 2);
  const int zoff = c * k - zlo;
  double ua = 0.0, ub = 0.0;
  for (int tt = 0; tt < tn; ++tt) {
    const int t = t0 + tt;
    const long long at = wbase + (long long)tt * D.ldM + m;
    const double* z = sm + tt * zw + zoff;
    double uu = 0.0;
#pragma unroll
    for (int m2 = 0; m2 < KMAX; ++m2)
      if (m2 < k) uu = fma(lr[m2], z[m2], uu);
    const double ystar = log(uu * uu + .0001);
    const double h = P.H[at];
    if (!(t & 1) || tt == 0) rng.uniform_pair(ST_SV_MIX_U, (uint32_t)i * nh + ((uint32_t)t >> 1), 0, ua, ub);
    const double uni = (t & 1) ? ub : ua;
    const double r = ystar - h;
    // The indicator only needs to know on which side of each cumulative weight u * tot falls.  A single-precision pass
    // (7 MUFU exponentials instead of 7 fp64 ones: the fp64 pipe was this kernel's bound) decides that whenever u * tot is
    // further than MIX_MARGIN * tot from every boundary - its weights are within 2e-5 relative of the fp64 ones (the
    // exponent a is formed in fp64 and rounded once: 2^-24 |a|, |a| < 88; plus ex2.approx) - and only the rare ambiguous element (~0.1 %), or one whose
    // weights underflow in fp32, takes the fp64 evaluation below.  The result is the fp64 decision in every case.
    int s;
    bool decided = false;
    {
      const float uf = (float)uni;
      float pf[7], totf = 0.f;
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        const double e = r - c_ksc_mm[cc];
        pf[cc] = s_lwf[cc] * __expf((float)(-e * e * s_h2v[cc]));  // exponent formed in fp64: its fp32 rounding is < 88 * 2^-24
        totf += pf[cc];
      }
      const float tgt = uf * totf, mar = MIX_MARGIN * totf;
      float cumf = 0.f;
      int cnt = 0;
      bool clear = totf > 1e-30f;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {  // the last boundary (cum = tot) is never ambiguous: u < 1
        cumf += pf[cc];
        clear = clear && fabsf(tgt - cumf) > mar;
        if (tgt < cumf) ++cnt;
      }
      s = 6 - cnt;  // = 7 - (cnt + 1): the seventh cumulative weight always exceeds u * tot
      decided = clear;
    }
    if (!decided) {
      double pdf[7], tot = 0.0;
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        const double e = r - c_ksc_mm[cc];
        pdf[cc] = s_lw[cc] * exp(-e * e * s_h2v[cc]);
        tot += pdf[cc];
      }
      const double target = uni * tot;  // u < cum / tot  <=>  u * tot < cum
      double cum = 0.0;
      int cnt = 0;
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        cum += pdf[cc];
        if (target < cum) ++cnt;
      }
      s = 7 - cnt;  // draw.h:468
      if (s > 6) s = 6;
    }
    const double iv = s_iv[s];
    P.Dinv[at] = iv;
    P.YLD[at] = (ystar - c_ksc_mm[s]) * iv;
  }
}

// The n standard normals per series of the log-volatility draw (draw.h:479-481; Philox stage ST_SV_H_Z, pair p of
// series i = normals 2p and 2p + 1), drawn by one thread per (series, pair) into the [t][m] buffer Zn.  They depend on
// nothing but the sweep counter, so the kernel runs on a side branch of the graph beside the Gram contraction instead of
// inside the one-thread-per-series recurrence of sv_solve_kernel, whose serial loop then only loads them.
__global__ void __launch_bounds__(256) sv_normals_kernel(const Dims D, const DevPtrs P) {
  const int w = blockIdx.z;
  const int n = P.win_n[w];
  const int pr = blockIdx.y, t = 2 * pr;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || m >= D.M) return;
  const int c = m / D.k, i = m - c * D.k;
  const int u = w * D.C + c;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const uint32_t nh = (uint32_t)((n + 1) / 2);
  double z0, z1;
  rng.normal_pair(ST_SV_H_Z, (uint32_t)i * nh + (uint32_t)pr, 0, z0, z1);
  const long long at = P.win_off[w] + (long long)t * D.ldM + m;
  P.Zn[at] = z0;
  if (t + 1 < n) P.Zn[at + D.ldM] = z1;
}

// h_i ~ N(K^-1 b, K^-1) with the tridiagonal K = H'H / sig + diag(1/v_s) (draw.h:469-487): one thread per
// series runs the bidiagonal Cholesky forward, then the combined mean + noise back substitution; accesses are
// [t][m], coalesced.  The three recurrences are rewritten so that each step of the dependent chain is ONE FMA:
//   pivots   q_t = diag_t - isig^2 / q_{t-1}  as the ratio N_t / D_t of the linear recurrence
//            N_t = diag_t N_{t-1} - isig^2 D_{t-1},  D_t = N_{t-1}   (re-normalised every SVS_B steps);
//            the divisions and w_t = 1 / sqrt(q_t) are then independent per t and pipeline freely;
//   forward  y_t = alpha_t + beta_t y_{t-1},   alpha_t = b_t w_t,  beta_t = isig w_{t-1} w_t;
//   backward x_t = gamma_t + delta_t x_{t+1},  gamma_t = (y_t + z_t) w_t,  delta_t = isig w_t^2.
constexpr int SVS_B = 16;  // time steps per block: 3 x 16 independent loads in flight per thread (the kernel runs ~140 threads per SM: latency-bound)
__global__ void __launch_bounds__(128) sv_solve_kernel(const Dims D, const DevPtrs P) {
  const int w = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= D.M) return;
  const int n = P.win_n[w];
  const long long off = P.win_off[w] + m;
  const int c = m / D.k, i = m % D.k;
  const int u = w * D.C + c;
  const double isig = 1.0 / P.lvol_sig[(size_t)u * D.k + i];
  const double isig2 = isig * isig;
  const double h0 = P.lvol_init[(size_t)u * D.k + i];
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const uint32_t nh = (uint32_t)((n + 1) / 2);
  double Nc = 1.0, Dc = 0.0;     // q_{-1} "= infinity": the first pivot has no sub-diagonal term
  double wprev = 0.0, yprev = 0.0;
  double dgn[SVS_B], bbn[SVS_B];  // the NEXT block's inputs: loaded before this block's stores are issued
#pragma unroll
  for (int r = 0; r < SVS_B; ++r) {
    const long long at = off + (long long)r * D.ldM;
    dgn[r] = r < n ? P.Dinv[at] : 1.0;
    bbn[r] = r < n ? P.YLD[at] : 0.0;
  }
  for (int t0 = 0; t0 < n; t0 += SVS_B) {
    double dg[SVS_B], bb[SVS_B], Nr[SVS_B], Dr[SVS_B];
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {
      dg[r] = dgn[r];
      bb[r] = bbn[r];
    }
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {
      const int t = t0 + SVS_B + r;
      const long long at = off + (long long)t * D.ldM;
      dgn[r] = t < n ? P.Dinv[at] : 1.0;
      bbn[r] = t < n ? P.YLD[at] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {  // chain: one FMA per step
      const int t = t0 + r;
      const double diag = (t < n - 1 ? 2.0 : 1.0) * isig + dg[r];
      const double Nn = fma(diag, Nc, -isig2 * Dc);
      Dc = Nc;
      Nc = Nn;
      Nr[r] = Nn;
      Dr[r] = Dc;
    }
    double wv[SVS_B];
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {  // w = 1 / sqrt(N / D), without the division unless N D overflows; independent: pipelined
      const double nd = Nr[r] * Dr[r];
      wv[r] = nd < 1e300 ? Dr[r] * rsqrt(nd) : rsqrt(Nr[r] / Dr[r]);
    }
    Nc = Nr[SVS_B - 1] / Dr[SVS_B - 1];  // re-normalise (q itself) for the next block
    Dc = 1.0;
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {
      const int t = t0 + r;
      if (t < n) {
        const double b = (t == 0 ? h0 * isig : 0.0) + bb[r];
        const double beta = isig * wprev * wv[r];
        const double y = fma(beta, yprev, b * wv[r]);
        const long long at = off + (long long)t * D.ldM;
        P.Dinv[at] = wv[r];
        P.YLD[at] = y;
        wprev = wv[r];
        yprev = y;
      }
    }
  }
  double next = 0.0, ssq = 0.0;
  const int nb = (n + SVS_B - 1) / SVS_B;
  double wvn[SVS_B], yvn[SVS_B], zvn[SVS_B];  // z: the standard normals of this sweep, pre-drawn by sv_normals_kernel
#pragma unroll
  for (int r = 0; r < SVS_B; ++r) {
    const int t = (nb - 1) * SVS_B + r;
    const long long at = off + (long long)t * D.ldM;
    wvn[r] = t < n ? P.Dinv[at] : 0.0;
    yvn[r] = t < n ? P.YLD[at] : 0.0;
    zvn[r] = t < n ? P.Zn[at] : 0.0;
  }
  for (int blk = nb - 1; blk >= 0; --blk) {
    const int t0 = blk * SVS_B;  // SVS_B is even, so Philox pairs (t even, t odd) never straddle blocks
    double wv[SVS_B], yv[SVS_B], zv[SVS_B];
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {
      wv[r] = wvn[r];
      yv[r] = yvn[r];
      zv[r] = zvn[r];
    }
    if (blk > 0) {
#pragma unroll
      for (int r = 0; r < SVS_B; ++r) {  // blocks below the last are full
        const long long at = off + (long long)(t0 - SVS_B + r) * D.ldM;
        wvn[r] = P.Dinv[at];
        yvn[r] = P.YLD[at];
        zvn[r] = P.Zn[at];
      }
    }
    double gam[SVS_B], del[SVS_B];
#pragma unroll
    for (int r = 0; r < SVS_B; ++r) {
      gam[r] = (yv[r] + zv[r]) * wv[r];
      del[r] = (t0 + r < n - 1) ? isig * wv[r] * wv[r] : 0.0;
    }
#pragma unroll
    for (int r = SVS_B - 1; r >= 0; --r) {
      const int t = t0 + r;
      if (t < n) {
        const double hnext = next;
        next = fma(del[r], next, gam[r]);
        P.H[off + (long long)t * D.ldM] = next;
        if (t < n - 1) ssq = fma(hnext - next, hnext - next, ssq);
      }
    }
  }
  ssq = fma(next - h0, next - h0, ssq);  // t = 0 against the initial state (draw.h:502-506)
  P.znorm[(size_t)D.U * D.k + (size_t)u * D.k + i] = ssq;  // per-series sum of squared increments, consumed by sv_finish
}

// (varsv_sigh / varsv_h0: sv_finish_kernel, kernels_sv_finish.cu)

// ------------------------------------------------------------------ warp-tile forms of updateSv / the mixture step
// Both elementwise SV kernels start from a small per-unit matrix product over the series index: (Y L')_{t,i} for
// updateSv, (Z L')_{t,i} for the mixture indicators.  One warp owns an 8 (time) x k (series) tile of one unit and forms
// it on the fp64 tensor pipe: DMMA m8n8k4 with A = the tile of Y / Z (rows = time, read straight from global memory:
// every 32-byte sector fetched is fully used), B = L' fragments from a shared-memory copy of the unit's L (k-steps above
// the diagonal block are skipped: L is lower triangular).  The accumulator fragment leaves every lane with the two
// neighbouring series (8 nt + 2 fc, + 1) of time point t0 + fr, on which the elementwise work then runs with two
// independent dependency chains per lane and no per-thread copy of a row of L (the thread-per-series versions kept 24-128
// doubles of L in registers: 128 registers, 25 % occupancy, latency-bound at 0.58 / 0.26 ms for C3).
constexpr int SVW_WARPS = 8;
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// the fp64 evaluation of the indicator (draw.h:458-468), out of line: it is the rare path and must not cost the common
// one its registers
__device__ __noinline__ int ksc_indicator_fp64(double r, double uni, const double* lw, const double* h2v) {
  double pdf[7], tot = 0.0;
#pragma unroll
  for (int cc = 0; cc < 7; ++cc) {
    const double e = r - c_ksc_mm[cc];
    pdf[cc] = lw[cc] * exp(-e * e * h2v[cc]);
    tot += pdf[cc];
  }
  const double target = uni * tot;  // u < cum / tot  <=>  u * tot < cum
  double cum = 0.0;
  int cnt = 0;
#pragma unroll
  for (int cc = 0; cc < 7; ++cc) {
    cum += pdf[cc];
    if (target < cum) ++cnt;
  }
  const int s = 7 - cnt;  // draw.h:468
  return s > 6 ? 6 : s;
}
// KS = ceil(k / 4) rounded up to the template bucket; MIX = false: updateSv (A = Y), true: mixture indicators (A = Z)
template <int KS, bool MIX>
__global__ void __launch_bounds__(SVW_WARPS * 32, 3) sv_tile_kernel(const Dims D, const DevPtrs P, const int tiles_per_cta) {
  extern __shared__ __align__(16) double sm[];  // L of the unit, row pitch ldl (odd)
  __shared__ double s_h2v[7], s_iv[7];
  __shared__ float s_lwf[7];
  __shared__ double s_lw[7];
  const int u = blockIdx.x, w = u / D.C, c = u - w * D.C;
  const int n = P.win_n[w], row0 = P.win_row0[w];
  const int k = D.k, ldl = k | 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fc = lane & 3;
  for (int e = tid; e < k * k; e += blockDim.x) {
    const int i = e / k, m = e - i * k;
    sm[i * ldl + m] = P.L[(size_t)u * k * k + e];
  }
  if (MIX && tid < 7) {
    const double v = c_ksc_v[tid];
    s_h2v[tid] = 0.5 / v;
    s_iv[tid] = 1.0 / v;
    s_lw[tid] = c_ksc_p[tid] / (sqrt(v) * sqrt(2 * 3.14159265358979323846));
    s_lwf[tid] = (float)s_lw[tid];
  }
  __syncthreads();
  const int ntt = (n + 7) >> 3;
  const int tile_lo = blockIdx.y * tiles_per_cta, tile_hi = min(ntt, tile_lo + tiles_per_cta);
  const long long ubase = P.win_off[w] + (long long)c * k;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const uint32_t nh = (uint32_t)((n + 1) / 2);
  const int nnt = (k + 7) >> 3;
  const bool vec2 = ((k & 1) == 0);  // series pairs are 16-byte aligned in the [t][m] arrays (ldM and the window offsets are even)
  for (int tile = tile_lo + warp; tile < tile_hi; tile += SVW_WARPS) {
    const int t = tile * 8 + fr;
    const bool tok = t < n;
    double af[KS];
    {
      const double* arow = MIX ? P.Z + ubase + (long long)t * D.ldM : P.Y + (size_t)(row0 + t) * k;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int m = 4 * ks + fc;
        af[ks] = (tok && m < k) ? __ldg(arow + m) : 0.0;
      }
    }
    for (int nt = 0; nt < nnt; ++nt) {
      double a0 = 0.0, a1 = 0.0;
      const int brow = min(8 * nt + fr, k - 1);
      const double* bl = sm + brow * ldl;
      const bool bok = 8 * nt + fr < k;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (4 * ks <= 8 * nt + 7) {  // warp-uniform: columns beyond the diagonal block of L are zero
          const int m = 4 * ks + fc;
          const double b = (bok && m < k) ? bl[m] : 0.0;
          dmma884(a0, a1, af[ks], b);
        }
      }
      const int i0 = 8 * nt + 2 * fc;  // this lane's series i0, i0 + 1 at time t
      const long long at = ubase + (long long)t * D.ldM + i0;
      if (!MIX) {
        // updateSv (tri.h:409): Dinv = exp(-h), YLD = (Y L')_{t,i} * Dinv
        if (tok && i0 < k) {
          if (vec2 && i0 + 1 < k) {
            const double2 h = *reinterpret_cast<const double2*>(P.H + at);
            const double d0 = exp(-h.x), d1 = exp(-h.y);
            *reinterpret_cast<double2*>(P.Dinv + at) = make_double2(d0, d1);
            *reinterpret_cast<double2*>(P.YLD + at) = make_double2(a0 * d0, a1 * d1);
          } else {
            const double d0 = exp(-P.H[at]);
            P.Dinv[at] = d0; P.YLD[at] = a0 * d0;
            if (i0 + 1 < k) { const double d1 = exp(-P.H[at + 1]); P.Dinv[at + 1] = d1; P.YLD[at + 1] = a1 * d1; }
          }
        }
        continue;
      }
      // ---- mixture indicators (draw.h:458-468).  Uniforms: Philox pair (series i, t >> 1) holds times 2p, 2p + 1; the lane
      // draws the pair of series i0 + (fr & 1) and swaps the half it does not need with the lane of the other parity of fr
      double ua, ub;
      {
        const int sc = min(i0 + (fr & 1), k - 1);
        rng.uniform_pair(ST_SV_MIX_U, (uint32_t)sc * nh + ((uint32_t)t >> 1), 0, ua, ub);
      }
      const double send = (fr & 1) ? ua : ub;
      const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
      const double un0 = (fr & 1) ? recv : ua;   // series i0:     even t -> own ua, odd t -> partner's ub
      const double un1 = (fr & 1) ? ub : recv;   // series i0 + 1: even t -> partner's ua, odd t -> own ub
      double hh0 = 0.0, hh1 = 0.0;
      if (tok && i0 < k) {
        if (vec2 && i0 + 1 < k) { const double2 h = *reinterpret_cast<const double2*>(P.H + at); hh0 = h.x; hh1 = h.y; }
        else { hh0 = P.H[at]; if (i0 + 1 < k) hh1 = P.H[at + 1]; }
      }
      double o_iv[2], o_b[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const double uu = h2 ? a1 : a0, hcur = h2 ? hh1 : hh0, uni = h2 ? un1 : un0;
        const double ystar = log(uu * uu + .0001);  // tri.h:401-402
        const double r = ystar - hcur;
        // The indicator only needs the side of u * tot relative to each cumulative weight.  A single-precision pass (exponent
        // formed in fp64 and rounded once, 7 MUFU exponentials: weights within 2e-5 relative of the fp64 ones) decides it
        // whenever u * tot is further than MIX_MARGIN * tot from every boundary; the rare ambiguous element (~0.1 %), or one
        // whose weights underflow in fp32, takes the fp64 evaluation.  The result is the fp64 decision in every case.
        int sidx;
        bool decided;
        {
          float pf[7], totf = 0.f;
#pragma unroll
          for (int cc = 0; cc < 7; ++cc) {
            const double e = r - c_ksc_mm[cc];
            pf[cc] = s_lwf[cc] * __expf((float)(-e * e * s_h2v[cc]));
            totf += pf[cc];
          }
          const float tgt = (float)uni * totf, mar = MIX_MARGIN * totf;
          float cumf = 0.f;
          int cnt = 0;
          bool clear = totf > 1e-30f;
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {  // the last boundary (cum = tot) is never ambiguous: u < 1
            cumf += pf[cc];
            clear = clear && fabsf(tgt - cumf) > mar;
            if (tgt < cumf) ++cnt;
          }
          sidx = 6 - cnt;
          decided = clear;
        }
        if (!decided) sidx = ksc_indicator_fp64(r, uni, s_lw, s_h2v);
        o_iv[h2] = s_iv[sidx];
        o_b[h2] = (ystar - c_ksc_mm[sidx]) * s_iv[sidx];
      }
      if (tok && i0 < k) {  // tridiagonal system's diagonal term 1 / v_s and right-hand-side term (y* - m_s) / v_s
        if (vec2 && i0 + 1 < k) {
          *reinterpret_cast<double2*>(P.Dinv + at) = make_double2(o_iv[0], o_iv[1]);
          *reinterpret_cast<double2*>(P.YLD + at) = make_double2(o_b[0], o_b[1]);
        } else {
          P.Dinv[at] = o_iv[0]; P.YLD[at] = o_b[0];
          if (i0 + 1 < k) { P.Dinv[at + 1] = o_iv[1]; P.YLD[at + 1] = o_b[1]; }
        }
      }
    }
  }
}
template <bool MIX>
static cudaError_t launch_sv_tile(const Dims& D, const DevPtrs& P, cudaStream_t st) {
  const int k = D.k, ntt = (D.nmax + 7) / 8;
  // enough CTAs to fill the machine a few times over, each warp keeping >= 2 time tiles
  int ysplit = std::max(1, std::min((ntt + 2 * SVW_WARPS - 1) / (2 * SVW_WARPS), (148 * 8 + D.U - 1) / D.U));
  const int tiles_per_cta = (ntt + ysplit - 1) / ysplit;
  ysplit = (ntt + tiles_per_cta - 1) / tiles_per_cta;
  const dim3 grid(D.U, ysplit);
  const size_t smem = (size_t)k * (k | 1) * sizeof(double);
#define BV_SVT(KSV)                                                                                             \
  do {                                                                                                          \
    if (smem > 48 * 1024) cudaFuncSetAttribute(sv_tile_kernel<KSV, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    sv_tile_kernel<KSV, MIX><<<grid, SVW_WARPS * 32, smem, st>>>(D, P, tiles_per_cta);                            \
  } while (0)
  const int ks = (k + 3) / 4;
  if (ks <= 2) BV_SVT(2);
  else if (ks <= 4) BV_SVT(4);
  else if (ks <= 5) BV_SVT(5);
  else if (ks <= 6) BV_SVT(6);
  else if (ks <= 8) BV_SVT(8);
  else if (ks <= 13) BV_SVT(13);
  else if (ks <= 16) BV_SVT(16);
  else if (ks <= 26) BV_SVT(26);
  else BV_SVT(32);
#undef BV_SVT
  return cudaGetLastError();
}

// ------------------------------------------------------------------ records
// updateRecords: row `rec_row` of every record matrix <- current state (config.h:850-857, 955-965,
// 796-804 and the prior records config.h:1071-1076, 1122-1125, 1183-1188, 1229-1233).
__global__ void __launch_bounds__(256) record_kernel(const Dims D, const DevPtrs P, const RecPtrs R) {
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int row = *P.rec_row;
  if (row >= R.cap) return;
  const int k = D.k, d = D.d, N = D.N, q = D.q, G = D.G, nrow = D.nrow;
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t ur = (size_t)u * R.cap + row;
  if (R.alpha) for (int e = tid; e < N; e += nt) R.alpha[ur * N + e] = P.alpha[(size_t)u * N + e];
  if (R.c) for (int e = tid; e < k; e += nt) R.c[ur * k + e] = P.cvec[(size_t)u * k + e];
  if (R.a) for (int e = tid; e < q; e += nt) R.a[ur * q + e] = P.contem[(size_t)u * q + e];
  if (R.d) for (int e = tid; e < k; e += nt) R.d[ur * k + e] = P.diag[(size_t)u * k + e];
  if (R.h0) for (int e = tid; e < k; e += nt) R.h0[ur * k + e] = P.lvol_init[(size_t)u * k + e];
  if (R.sigh) for (int e = tid; e < k; e += nt) R.sigh[ur * k + e] = P.lvol_sig[(size_t)u * k + e];
  if (R.h || R.h_last) {
    const int n = P.win_n[w];
    const long long off = P.win_off[w] + (long long)c * k;
    if (R.h)  // lvol_draw.transpose().reshaped(): time-major, k contiguous (config.h:962)
      for (int e = tid; e < n * k; e += nt) R.h[ur * ((size_t)D.nmax * k) + e] = P.H[off + (long long)(e / k) * D.ldM + (e % k)];
    if (R.h_last)
      for (int e = tid; e < k; e += nt) R.h_last[ur * k + e] = P.H[off + (long long)(n - 1) * D.ldM + e];
  }
  if (R.alpha_sparse)
    for (int e = tid; e < N; e += nt) R.alpha_sparse[ur * N + e] = P.sparse_coef[(size_t)u * d * k + (e / nrow) * d + (e % nrow)];
  if (R.c_sparse) for (int e = tid; e < k; e += nt) R.c_sparse[ur * k + e] = P.sparse_coef[(size_t)u * d * k + e * d + nrow];
  if (R.a_sparse) for (int e = tid; e < q; e += nt) R.a_sparse[ur * q + e] = P.sparse_contem[(size_t)u * q + e];
  if (R.gamma) for (int e = tid; e < N; e += nt) R.gamma[ur * N + e] = P.hN0[(size_t)u * N + e];
  if (R.lambda) for (int e = tid; e < N; e += nt) R.lambda[ur * N + e] = P.hN0[(size_t)u * N + e];
  if (R.eta) for (int e = tid; e < G; e += nt) R.eta[ur * G + e] = P.hG0[(size_t)u * G + e];
  if (R.tau && tid == 0) R.tau[ur] = P.hs[(size_t)u * HS_SLOTS + HS_GLOBAL];
  if (R.kappa) for (int e = tid; e < N; e += nt) R.kappa[ur * N + e] = P.hN2[(size_t)u * N + e];
}
__global__ void bump_kernel(uint32_t* iter, int* rec_row, int bump_iter, int bump_rec) {
  if (bump_iter) *iter += 1u;
  if (bump_rec) *rec_row += 1;
}

// ------------------------------------------------------------------ launch wrappers
static size_t coef_smem(const Dims& D, int CH, bool vglobal, bool pglobal) {
  size_t n = CH + (size_t)D.k * D.k + 6 * (size_t)D.d + D.k;
  if (!pglobal) n += (D.Pp + 1) & ~1;
  if (!vglobal) n += 2 * (size_t)D.k * D.d;
  return n * sizeof(double);
}
bool coef_plan(const Dims& D, size_t smem_limit, int* CH_out, bool* vglobal_out, bool* pglobal_out) {
  int CH = D.Pp < 4096 ? D.Pp : 4096;
  if (CH < D.d) CH = D.d;
  bool vglobal = false, pglobal = false;
  if (coef_smem(D, CH, false, false) > smem_limit) vglobal = true;
  if (coef_smem(D, CH, vglobal, false) > smem_limit) pglobal = true;
  if (coef_smem(D, CH, vglobal, pglobal) > smem_limit) {
    CH = D.d;
    if (coef_smem(D, CH, vglobal, pglobal) > smem_limit) return false;
  }
  *CH_out = CH;
  *vglobal_out = vglobal;
  *pglobal_out = pglobal;
  return true;
}
cudaError_t launch_coef(const Dims& D, const DevPtrs& P, double* vscratch, double* pscratch, size_t smem_limit, cudaStream_t st) {
  int CH = 0;
  bool vglobal = false, pglobal = false;
  if (!coef_plan(D, smem_limit, &CH, &vglobal, &pglobal)) return cudaErrorInvalidConfiguration;
  if ((vglobal && !vscratch) || (pglobal && !pscratch)) return cudaErrorInvalidValue;
  const size_t smem = coef_smem(D, CH, vglobal, pglobal);
  if (D.sv) {
    cudaFuncSetAttribute(coef_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_kernel<true><<<D.U, 256, smem, st>>>(D, P, CH, vglobal ? vscratch : nullptr, pglobal ? pscratch : nullptr);
  } else {
    cudaFuncSetAttribute(coef_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_kernel<false><<<D.U, 256, smem, st>>>(D, P, CH, vglobal ? vscratch : nullptr, pglobal ? pscratch : nullptr);
  }
  return cudaGetLastError();
}
// Batched precision draw (building block exported for stage-level parity tests): one CTA per system.
__global__ void __launch_bounds__(256) draw_coef_batched_kernel(int d, const double* __restrict__ Pp, const double* __restrict__ b,
                                                                const double* __restrict__ z, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int np = d * (d + 1) / 2;
  double* Pk = sm;
  double* r = Pk + np;
  double* zz = r + d;
  double* x = zz + d;
  __shared__ int bad;
  const size_t s = blockIdx.x;
  for (int p = threadIdx.x; p < np; p += blockDim.x) Pk[p] = Pp[s * np + p];
  for (int a = threadIdx.x; a < d; a += blockDim.x) { r[a] = b[s * d + a]; zz[a] = z[s * d + a]; }
  if (threadIdx.x == 0) bad = 0;
  chol_packed_block(Pk, d, &bad);
  if (threadIdx.x < 32) precision_solve_warp(Pk, d, r, zz, x);
  __syncthreads();
  for (int a = threadIdx.x; a < d; a += blockDim.x) out[s * d + a] = x[a];
}
cudaError_t launch_draw_coef_batched(long long batch, int d, const double* Pp, const double* b, const double* z, double* out,
                                     cudaStream_t st) {
  const size_t smem = ((size_t)d * (d + 1) / 2 + 3 * d) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaFuncSetAttribute(draw_coef_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  draw_coef_batched_kernel<<<(unsigned)batch, 256, smem, st>>>(d, Pp, b, z, out);
  return cudaGetLastError();
}
// ---- large-design SV path: the weights of the direct precision build -------------------------------------------------
// P_j = X' diag(w_j) X with w_tj = sum_{i>=j} L_ij^2 / s_ti^2 (the weighted Gram of equation j's Kronecker design,
// triangular.h:286 + draw.h:380).  Where the S_i = X' D_i X themselves are not needed (designs with d > 128: the
// sequential part runs in the time domain, kernels_seq_big.cu) the Gram contraction G1 takes these weights as its A
// operand and writes the k precision matrices of a unit directly - no S_i, no [k x k] x [k x Pp] mixing pass over them.
// One CTA per (unit, 32 time points): L o L in shared memory, one thread per (t, j).
constexpr int WMX_TT = 32;
__global__ void __launch_bounds__(256) sv_wmix_kernel(const Dims D, const DevPtrs P, double* __restrict__ Wt) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C, k = D.k;
  const int n = P.win_n[w], t0 = blockIdx.y * WMX_TT;
  if (t0 >= n) return;
  double* L2 = sm;           // [k][k] L_ij^2
  double* dt = L2 + k * k;   // [TT][k]
  const long long off = P.win_off[w] + (long long)c * k;
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    const double l = P.L[(size_t)u * k * k + e];
    L2[e] = l * l;
  }
  for (int e = threadIdx.x; e < WMX_TT * k; e += blockDim.x) {
    const int tt = e / k, i = e - tt * k;
    dt[e] = t0 + tt < n ? P.Dinv[off + (long long)(t0 + tt) * D.ldM + i] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < WMX_TT * k; e += blockDim.x) {
    const int tt = e / k, j = e - tt * k;
    if (t0 + tt >= n) continue;
    double acc = 0.0;
    for (int i = j; i < k; ++i) acc = fma(L2[i * k + j], dt[tt * k + i], acc);
    Wt[off + (long long)(t0 + tt) * D.ldM + j] = acc;
  }
}
cudaError_t launch_sv_wmix(const Dims& D, const DevPtrs& P, double* Wt, cudaStream_t st) {
  const size_t smem = ((size_t)D.k * D.k + (size_t)WMX_TT * D.k) * sizeof(double);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaFuncSetAttribute(sv_wmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sv_wmix_kernel<<<dim3(D.U, (D.nmax + WMX_TT - 1) / WMX_TT), 256, smem, st>>>(D, P, Wt);
  return cudaGetLastError();
}
cudaError_t launch_sv_prep(const Dims& D, const DevPtrs& P, cudaStream_t st) {
  if (D.k <= 128 && getenv("BVHAR_SV_THREAD_KERNELS") == nullptr) return launch_sv_tile<false>(D, P, st);
  const dim3 grid((D.M + 255) / 256, (D.nmax + SVP_TT - 1) / SVP_TT, D.W);
  const size_t smem = (size_t)SVP_TT * D.k * sizeof(double);
  if (D.k <= 8) sv_prep_kernel<8><<<grid, 256, smem, st>>>(D, P);
  else if (D.k <= 16) sv_prep_kernel<16><<<grid, 256, smem, st>>>(D, P);
  else if (D.k <= 24) sv_prep_kernel<24><<<grid, 256, smem, st>>>(D, P);
  else if (D.k <= 32) sv_prep_kernel<32><<<grid, 256, smem, st>>>(D, P);
  else if (D.k <= 64) sv_prep_kernel<64><<<grid, 256, smem, st>>>(D, P);
  else sv_prep_kernel<128><<<grid, 256, smem, st>>>(D, P);
  return cudaGetLastError();
}
cudaError_t launch_sv_state(int part, const Dims& D, const DevPtrs& P, cudaStream_t st) {
  if (part == 3) {  // the pre-drawn normals of sv_solve
    sv_normals_kernel<<<dim3((D.M + 255) / 256, (D.nmax + 1) / 2, D.W), 256, 0, st>>>(D, P);
    return cudaGetLastError();
  }
  if (part == 0 && D.k <= 128 && getenv("BVHAR_SV_THREAD_KERNELS") == nullptr) return launch_sv_tile<true>(D, P, st);
  if (part == 0) {
    const dim3 grid((D.M + 255) / 256, (D.nmax + SVP_TT - 1) / SVP_TT, D.W);
    const size_t smem = (size_t)SVP_TT * (256 + 2 * D.k) * sizeof(double);
    if (D.k <= 8) sv_mix_kernel<8><<<grid, 256, smem, st>>>(D, P);
    else if (D.k <= 16) sv_mix_kernel<16><<<grid, 256, smem, st>>>(D, P);
    else if (D.k <= 24) sv_mix_kernel<24><<<grid, 256, smem, st>>>(D, P);
    else if (D.k <= 32) sv_mix_kernel<32><<<grid, 256, smem, st>>>(D, P);
    else if (D.k <= 64) {
      cudaFuncSetAttribute(sv_mix_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      sv_mix_kernel<64><<<grid, 256, smem, st>>>(D, P);
    } else {
      cudaFuncSetAttribute(sv_mix_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      sv_mix_kernel<128><<<grid, 256, smem, st>>>(D, P);
    }
  } else if (part == 1) {
    sv_solve_kernel<<<dim3((D.M + 127) / 128, D.W), 128, 0, st>>>(D, P);
  } else {
    return launch_sv_finish(D, P, st);
  }
  return cudaGetLastError();
}
cudaError_t launch_record(const Dims& D, const DevPtrs& P, const RecPtrs& R, cudaStream_t st) {
  record_kernel<<<D.U, 256, 0, st>>>(D, P, R);
  return cudaGetLastError();
}
cudaError_t launch_bump(const DevPtrs& P, int bump_iter, int bump_rec, cudaStream_t st) {
  bump_kernel<<<1, 1, 0, st>>>(P.iter, P.rec_row, bump_iter, bump_rec);
  return cudaGetLastError();
}

}  // namespace bvhar
