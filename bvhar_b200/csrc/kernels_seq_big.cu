// kernels_seq_big.cu - the sequential part of the coefficient stage for large designs (see kernels_chol_big.cu);
// compiled once per BV_NC = 8 | 16 (ceil(d / 32) class).
#include <algorithm>

#include "block_linalg.cuh"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "warp_linalg.cuh"

#ifndef BV_NC
#error "compile with -DBV_NC=8|16"
#endif
#define BV_CAT2(a, b) a##b
#define BV_CAT(a, b) BV_CAT2(a, b)

namespace bvhar {

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_PB = 36;  // pitch of the 32 x 32 B tile: 36 = 4 (mod 16) => conflict-free DMMA fragment loads

__device__ __forceinline__ size_t rowoff(int a) { return (size_t)a * (a + 1) / 2; }

// ---- blocked two-sided triangular solve by the whole CTA, factor in global memory ---------------------------
// delta = L^-T (L^-1 r + z).  y (shared, d doubles) holds r on entry and delta on exit; z in shared memory.
// The (inverted) diagonal block of step jb +- 1 is fetched into registers while step jb's off-diagonal update runs, so the
// 2 ceil(d / 32) dependent block steps do not each start with an exposed global-memory round trip; blockDim.x >= 256.
__device__ void precision_solve_cta(const double* __restrict__ Lg, int d, double* y, const double* z, double* blk /* 32 x 33 */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = blockDim.x;  // CB_THREADS (256) or TD_THREADS (1024)
  constexpr int DPT = 4;      // diagonal-block elements per thread at 256 threads (fewer are used by larger blocks)
  const int nb = (d + 31) / 32;
  double dv[DPT];
  auto load_diag = [&](int jb) {
    const int c0 = 32 * jb, nc = min(32, d - c0);
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      const int e = tid + q * nt, rr = e >> 5, cc = e & 31;
      dv[q] = (e < 1024 && rr < nc && cc <= rr) ? Lg[rowoff(c0 + rr) + c0 + cc] : 0.0;
    }
  };
  auto store_diag = [&]() {
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      const int e = tid + q * nt;
      if (e < 1024) blk[(e >> 5) * 33 + (e & 31)] = dv[q];
    }
  };
  // forward: L y = r
  load_diag(0);
  for (int jb = 0; jb < nb; ++jb) {
    const int c0 = 32 * jb, nc = min(32, d - c0);
    __syncthreads();
    store_diag();
    __syncthreads();
    if (warp == 0) {  // y_blk = Linv_blk y_blk (the diagonal blocks hold the inverse, kernels_chol_big.cu)
      double v0 = 0.0, v1 = 0.0;
#pragma unroll 8
      for (int c = 0; c < 32; c += 2) {
        v0 = fma(blk[lane * 33 + c], c < nc ? y[c0 + c] : 0.0, v0);
        v1 = fma(blk[lane * 33 + c + 1], c + 1 < nc ? y[c0 + c + 1] : 0.0, v1);
      }
      __syncwarp();
      if (lane < nc) y[c0 + lane] = v0 + v1;
    }
    if (jb + 1 < nb) load_diag(jb + 1);
    __syncthreads();
    for (int r = c0 + 32 + tid; r < d; r += nt) {  // rows below: y_r -= L[r][c0 : c0+32] . y_blk
      const double* row = Lg + rowoff(r) + c0;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        s0 = fma(row[c], y[c0 + c], s0);
        s1 = fma(row[c + 1], y[c0 + c + 1], s1);
        s2 = fma(row[c + 2], y[c0 + c + 2], s2);
        s3 = fma(row[c + 3], y[c0 + c + 3], s3);
      }
      y[r] -= (s0 + s1) + (s2 + s3);
    }
  }
  load_diag(nb - 1);
  __syncthreads();
  for (int a = tid; a < d; a += nt) y[a] += z[a];
  // backward: L' x = y + z
  for (int jb = nb - 1; jb >= 0; --jb) {
    const int c0 = 32 * jb, nc = min(32, d - c0);
    __syncthreads();
    store_diag();
    __syncthreads();
    if (warp == 0) {  // x_blk = Linv_blk' y_blk
      double v0 = 0.0, v1 = 0.0;
#pragma unroll 8
      for (int r = 0; r < 32; r += 2) {
        v0 = fma(blk[r * 33 + lane], r < nc ? y[c0 + r] : 0.0, v0);
        v1 = fma(blk[(r + 1) * 33 + lane], r + 1 < nc ? y[c0 + r + 1] : 0.0, v1);
      }
      __syncwarp();
      if (lane < nc) y[c0 + lane] = v0 + v1;
    }
    if (jb > 0) load_diag(jb - 1);
    __syncthreads();
    for (int b = tid; b < c0; b += nt) {  // columns to the left: y_b -= sum_a L[c0 + a][b] x_a   (coalesced in b)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      if (nc == 32) {  // full block: all 32 loads of the thread in flight together
#pragma unroll
        for (int a = 0; a < 32; a += 4) {
          s0 = fma(Lg[rowoff(c0 + a) + b], y[c0 + a], s0);
          s1 = fma(Lg[rowoff(c0 + a + 1) + b], y[c0 + a + 1], s1);
          s2 = fma(Lg[rowoff(c0 + a + 2) + b], y[c0 + a + 2], s2);
          s3 = fma(Lg[rowoff(c0 + a + 3) + b], y[c0 + a + 3], s3);
        }
      } else {
        for (int a = 0; a < nc; ++a) s0 = fma(Lg[rowoff(c0 + a) + b], y[c0 + a], s0);
      }
      y[b] -= (s0 + s1) + (s2 + s3);
    }
  }
  __syncthreads();
}

// ---- B (large d): the sequential-in-j part, one CTA per unit ----------------------------------------------------
template <bool SV, int NC>
__global__ void __launch_bounds__(CB_THREADS, 2) coef_seq_big_kernel(const Dims D, const DevPtrs P, const double* __restrict__ Pfac,
                                                                   double* __restrict__ hscr, const double* __restrict__ Rres) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int k = D.k, d = D.d, ldM = D.ldM;
  double* Lsm = sm;                          // packed lower k(k+1)/2 (row i at i(i+1)/2): half the square, so that two CTAs share an SM
  double* yv = Lsm + (((size_t)k * (k + 1) / 2 + 1) & ~(size_t)1);  // [d] rhs -> delta
  double* zv = yv + d;                       // [d]
  double* tv = zv + d;                       // [d] LDLT: X'X delta
  double* blk = tv + d;                      // [32][33]
  double* wscr = blk + 32 * 33;              // [nw][d]
  double* idg = wscr + (size_t)nw * d;       // [k]
  double* h = hscr + (size_t)u * k * d;      // [k][d] C_i - S_i v_i (global, L2-resident)
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  double* Ac = P.Acoef + (size_t)w * d * ldM + (size_t)c * k;
  const double* prec_u = P.prec + (size_t)u * k * d;
  const double* xn = P.xnorm + (size_t)w * d;
  const double* X2 = SV ? nullptr : P.XtX + (size_t)w * D.ldS;

  for (int e = tid; e < k * k; e += blockDim.x) {
    const int i = e / k, m = e - i * k;
    if (m <= i) Lsm[i * (i + 1) / 2 + m] = P.L[(size_t)u * k * k + e];
  }
  if (!SV) for (int i = tid; i < k; i += blockDim.x) idg[i] = 1.0 / P.diag[(size_t)u * k + i];
  __syncthreads();
  if (!SV && Rres) {
    // LDLT with the residual contraction: h_i = (X'Y - X'X A) L_i' / d_i = R L_i' / d_i, no symmetric product (the k products
    // with the same X'X per unit were a third of this kernel's L2 traffic)
    const double* Ru = Rres + (size_t)w * d * ldM + (size_t)c * k;
    for (int i = warp; i < k; i += nw) {
      const double* Li = Lsm + i * (i + 1) / 2;
      for (int a = lane; a < d; a += 32) {
        const double* ra = Ru + (size_t)a * ldM;
        double s0 = 0.0, s1 = 0.0;
        int m = 0;
        for (; m + 1 <= i; m += 2) { s0 = fma(Li[m], ra[m], s0); s1 = fma(Li[m + 1], ra[m + 1], s1); }
        if (m <= i) s0 = fma(Li[m], ra[m], s0);
        h[(size_t)i * d + a] = (s0 + s1) * idg[i];
      }
    }
  } else
  for (int i = warp; i < k; i += nw) {  // h_i = C_i - S_i v_i, v_i = A L_i'
    double* v = wscr + (size_t)warp * d;
    for (int a = lane; a < d; a += 32) {
      double sacc = 0.0;
      for (int m = 0; m <= i; ++m) sacc = fma(Lsm[i * (i + 1) / 2 + m], Ac[(size_t)a * ldM + m], sacc);
      v[a] = sacc;
    }
    __syncwarp();
    double out[NC];
    symv_packed_warp<NC>(SV ? P.S + ((size_t)u * k + i) * D.ldS : X2, v, d, out);
#pragma unroll
    for (int sidx = 0; sidx < NC; ++sidx) {
      const int b = lane + 32 * sidx;
      if (b < d) {
        if (SV) {
          h[(size_t)i * d + b] = P.Cm[((size_t)u * k + i) * D.ldX + b] - out[sidx];
        } else {
          const double* xty = P.XtY + ((size_t)w * d + b) * k;
          double sacc = 0.0;
          for (int m = 0; m <= i; ++m) sacc = fma(xty[m], Lsm[i * (i + 1) / 2 + m], sacc);
          h[(size_t)i * d + b] = (sacc - out[sidx]) * idg[i];
        }
      }
    }
    __syncwarp();
  }
  __threadfence_block();
  __syncthreads();

  for (int j = 0; j < k; ++j) {
    for (int a = tid; a < d; a += blockDim.x) {
      double acc = prec_u[j * d + a] * (P.prior_mean[j * d + a] - Ac[(size_t)a * ldM + j]);
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int i = j;
      for (; i + 3 < k; i += 4) {  // four independent loads / chains per step
        acc = fma(Lsm[i * (i + 1) / 2 + j], h[(size_t)i * d + a], acc);
        a1 = fma(Lsm[(i + 1) * (i + 2) / 2 + j], h[(size_t)(i + 1) * d + a], a1);
        a2 = fma(Lsm[(i + 2) * (i + 3) / 2 + j], h[(size_t)(i + 2) * d + a], a2);
        a3 = fma(Lsm[(i + 3) * (i + 4) / 2 + j], h[(size_t)(i + 3) * d + a], a3);
      }
      for (; i < k; ++i) acc = fma(Lsm[i * (i + 1) / 2 + j], h[(size_t)i * d + a], acc);
      acc = (acc + a1) + (a2 + a3);
      yv[a] = acc;
      zv[a] = rng.normal_elem(ST_COEF_Z, (uint32_t)(j * d + a));
    }
    precision_solve_cta(Pfac + ((size_t)u * k + j) * D.ldS, d, yv, zv, blk);  // starts and ends with __syncthreads
    for (int a = tid; a < d; a += blockDim.x) {
      const double val = Ac[(size_t)a * ldM + j] + yv[a];
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
    if (j + 1 < k) {
      if (SV) {
        for (int i = j + 1 + warp; i < k; i += nw) {
          double out[NC];
          symv_packed_warp<NC>(P.S + ((size_t)u * k + i) * D.ldS, yv, d, out);
          const double lij = Lsm[i * (i + 1) / 2 + j];
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) h[(size_t)i * d + b] = fma(-lij, out[sidx], h[(size_t)i * d + b]);
          }
        }
      } else {
        // X'X delta once, rows split over the warps: warp q takes the 32-row blocks q, q + nw, ... and leaves its partial
        // product in its own row of wscr; the partials are then added in warp order (a shared-memory atomicAdd here made
        // the last bits of the draw depend on the warp scheduling: the trajectories were not reproducible)
        {
          double out[NC];
          symv_packed_warp<NC>(X2, yv, d, out, warp, nw);
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) wscr[(size_t)warp * d + b] = out[sidx];
          }
        }
        __syncthreads();
        for (int a = tid; a < d; a += blockDim.x) {
          double s = 0.0;
          for (int q = 0; q < nw; ++q) s += wscr[(size_t)q * d + a];
          tv[a] = s;
        }
        __syncthreads();
        for (int e = tid; e < (k - j - 1) * d; e += blockDim.x) {
          const int i = j + 1 + e / d, a = e % d;
          h[(size_t)i * d + a] = fma(-Lsm[i * (i + 1) / 2 + j] * idg[i], tv[a], h[(size_t)i * d + a]);
        }
      }
    }
    __threadfence_block();
    __syncthreads();
  }
}


#if BV_NC == 8
// ---- B' (large d, SV): the sequential-in-j part in the TIME domain, one CTA per unit ---------------------------
// The Gram-domain recurrence above applies every draw delta_j to all h_i = X' D_i e_i (i > j): (k - j) packed
// symmetric matrix-vector products per equation, k^2/2 * d^2/2 doubles of S_i traffic per unit and sweep (200 MB
// at C4, k = 50, d = 201: HBM-bound).  Here the residual is carried instead of its Gram image:
//     E = (Y - X A) L'   [n x k]   (kept transposed, series-major, next to D' = Dinv' in a per-unit scratch)
//     b_j     = prec_j o (mean_j - a_j) + X' u_j,     u_j,t = sum_{i>=j} L_ij E_ti / s_ti^2      (tri.h:286-299)
//     delta_j = P_j^-1 b_j + chol(P_j)^-T z           (draw.h:372-383; the factor comes from coef_chol_cta_kernel)
//     E_ti   -= L_ij (X delta_j)_t   for i >= j
// i.e. two passes over the window's X (L2-resident, shared by every chain and every overlapping window) and one
// pass over the unit's own 2 n (k - j) doubles per equation.  Same arithmetic as update_coef_efficient in the oracle
// up to the order of summation.
// FRES: the packed factor of equation j + 1 is prefetched into shared memory with cp.async while equation j's
// matrix-vector passes run, so that the 2 x ceil(d/32) dependent block steps of the solve never wait on HBM.
constexpr int TD_THREADS = 1024;
constexpr int TD_TCH = 32;  // time points per prologue tile

// delta = L^-T (L^-1 r + z) with the factor in shared memory (packed lower, diagonal blocks inverted).
__device__ void precision_solve_smem(const double* __restrict__ Lf, int d, double* y, const double* z) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int nb = (d + 31) / 32;
  for (int jb = 0; jb < nb; ++jb) {  // forward: L y = r
    const int c0 = 32 * jb, c1 = c0 + 32, nc = min(32, d - c0);
    __syncthreads();
    if (warp == 0) {  // y_blk = Linv_blk y_blk
      double v0 = 0.0, v1 = 0.0;
      if (lane < nc) {
        const double* lr = Lf + rowoff(c0 + lane) + c0;
        int c = 0;
        for (; c + 1 <= lane; c += 2) { v0 = fma(lr[c], y[c0 + c], v0); v1 = fma(lr[c + 1], y[c0 + c + 1], v1); }
        if (c <= lane) v0 = fma(lr[c], y[c0 + c], v0);
      }
      __syncwarp();
      if (lane < nc) y[c0 + lane] = v0 + v1;
    }
    __syncthreads();
    for (int base = c1; base < d; base += nt / 4) {  // rows below, four threads per row (eight columns each)
      const int r = base + (tid >> 2), p8 = (tid & 3) * 8;
      double s0 = 0.0, s1 = 0.0;
      if (r < d) {
        const double* row = Lf + rowoff(r) + c0 + p8;
        const double* yy = y + c0 + p8;
#pragma unroll
        for (int c = 0; c < 8; c += 2) { s0 = fma(row[c], yy[c], s0); s1 = fma(row[c + 1], yy[c + 1], s1); }
      }
      s0 += s1;
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      if (r < d && (tid & 3) == 0) y[r] -= s0;
    }
  }
  __syncthreads();
  for (int a = tid; a < d; a += nt) y[a] += z[a];
  for (int jb = nb - 1; jb >= 0; --jb) {  // backward: L' x = y + z
    const int c0 = 32 * jb, nc = min(32, d - c0);
    __syncthreads();
    if (warp == 0) {  // x_blk = Linv_blk' y_blk
      double v0 = 0.0, v1 = 0.0;
      if (lane < nc) {
        int r = lane;
        for (; r + 1 < nc; r += 2) {
          v0 = fma(Lf[rowoff(c0 + r) + c0 + lane], y[c0 + r], v0);
          v1 = fma(Lf[rowoff(c0 + r + 1) + c0 + lane], y[c0 + r + 1], v1);
        }
        if (r < nc) v0 = fma(Lf[rowoff(c0 + r) + c0 + lane], y[c0 + r], v0);
      }
      __syncwarp();
      if (lane < nc) y[c0 + lane] = v0 + v1;
    }
    __syncthreads();
    for (int b = tid; b < c0; b += nt) {  // columns to the left: y_b -= sum_a L[c0 + a][b] x_a
      double s0 = 0.0, s1 = 0.0;
      int a = 0;
      for (; a + 1 < nc; a += 2) {
        s0 = fma(Lf[rowoff(c0 + a) + b], y[c0 + a], s0);
        s1 = fma(Lf[rowoff(c0 + a + 1) + b], y[c0 + a + 1], s1);
      }
      if (a < nc) s0 = fma(Lf[rowoff(c0 + a) + b], y[c0 + a], s0);
      y[b] -= s0 + s1;
    }
  }
  __syncthreads();
}

struct TdLayout {  // shared-memory carve-up (doubles), the same on host and device
  int astr, nsplit, tstr, nsa, part, big;
  size_t total;
};
__host__ __device__ inline TdLayout td_layout(int k, int d, int npad, int ldS, bool fres) {
  TdLayout L;
  L.astr = (d + 31) & ~31;
  L.nsplit = max(1, TD_THREADS / L.astr);
  L.tstr = (npad + 31) & ~31;
  L.nsa = max(1, TD_THREADS / L.tstr);
  L.part = max(L.nsplit * L.astr, L.nsa * L.tstr);
  const int tile = 2 * TD_TCH * (k + 1);
  L.big = fres ? max(ldS, tile) : tile;
  L.total = (((size_t)k * k + 2 * (size_t)d + npad + L.part + 32 * 33 + 1) & ~(size_t)1) + L.big;  // `big` 16-byte aligned (cp.async)
  return L;
}

template <bool FRES>
__global__ void __launch_bounds__(TD_THREADS) coef_seq_time_kernel(const Dims D, const DevPtrs P, const double* __restrict__ Pfac,
                                                                    double* __restrict__ escr, const int npad) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int k = D.k, d = D.d, ldM = D.ldM;
  const int n = P.win_n[w], row0 = P.win_row0[w];
  const TdLayout LY = td_layout(k, d, npad, D.ldS, FRES);
  const int astr = LY.astr, nsplit = LY.nsplit, tstr = LY.tstr, nsa = LY.nsa;
  double* Lsm = sm;                          // [k][k]
  double* yv = Lsm + (size_t)k * k;          // [d] rhs -> delta
  double* zv = yv + d;                       // [d]
  double* uv = zv + d;                       // [npad]
  double* part = uv + npad;                  // [nsplit][astr] | [nsa][tstr]
  double* blk = part + LY.part;              // [32][33] (global-factor solve only)
  double* big = sm + (LY.total - LY.big);    // prologue tiles, then (FRES) the packed factor; 16-byte aligned
  double* __restrict__ ET = escr + (size_t)u * 2 * k * npad;  // [k][npad]
  double* __restrict__ DT = ET + (size_t)k * npad;            // [k][npad]
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  double* Ac = P.Acoef + (size_t)w * d * ldM + (size_t)c * k;
  const double* prec_u = P.prec + (size_t)u * k * d;
  const double* xn = P.xnorm + (size_t)w * d;
  const double* Xw = P.X + (size_t)row0 * D.ldX;
  const double* XTw = P.XT + row0;
  const double* Fg = Pfac + (size_t)u * k * D.ldS;
  auto prefetch_factor = [&](int j) {  // whole packed factor of equation j -> shared memory, asynchronously
    const double* src = Fg + (size_t)j * D.ldS;
    for (int e = 2 * tid; e < D.ldS; e += 2 * nt) cp_async16(big + e, src + e, 16);
    cp_async_commit();
  };

  for (int e = tid; e < k * k; e += nt) Lsm[e] = P.L[(size_t)u * k * k + e];
  // prologue: E' and D' from the [t][m] arrays (latent_innov is Y - X A of the current coefficients, tri.h:342)
  {
    const int kp = k + 1;
    double* zs = big;
    double* ds = big + TD_TCH * kp;
    const double* Zg = P.Z + P.win_off[w] + (size_t)c * k;
    const double* Dg = P.Dinv + P.win_off[w] + (size_t)c * k;
    for (int t0 = 0; t0 < n; t0 += TD_TCH) {
      const int tn = min(TD_TCH, n - t0);
      __syncthreads();
      for (int e = tid; e < tn * k; e += nt) {
        const int tl = e / k, m = e - tl * k;
        zs[tl * kp + m] = Zg[(size_t)(t0 + tl) * ldM + m];
        ds[tl * kp + m] = Dg[(size_t)(t0 + tl) * ldM + m];
      }
      __syncthreads();
      for (int e = tid; e < TD_TCH * k; e += nt) {
        const int tl = e & (TD_TCH - 1), i = e / TD_TCH;
        if (tl < tn) {
          const double* lr = Lsm + i * k;
          const double* zr = zs + tl * kp;
          double s0 = 0.0, s1 = 0.0;
          int m = 0;
          for (; m + 1 <= i; m += 2) { s0 = fma(lr[m], zr[m], s0); s1 = fma(lr[m + 1], zr[m + 1], s1); }
          if (m <= i) s0 = fma(lr[m], zr[m], s0);
          ET[(size_t)i * npad + t0 + tl] = s0 + s1;
          DT[(size_t)i * npad + t0 + tl] = ds[tl * kp + i];
        }
      }
    }
  }
  __syncthreads();
  if (FRES) prefetch_factor(0);
  for (int t = tid; t < n; t += nt) {  // u_0
    double s0 = 0.0, s1 = 0.0;
    int i = 0;
    for (; i + 1 < k; i += 2) {
      s0 = fma(Lsm[i * k] * ET[(size_t)i * npad + t], DT[(size_t)i * npad + t], s0);
      s1 = fma(Lsm[(i + 1) * k] * ET[(size_t)(i + 1) * npad + t], DT[(size_t)(i + 1) * npad + t], s1);
    }
    if (i < k) s0 = fma(Lsm[i * k] * ET[(size_t)i * npad + t], DT[(size_t)i * npad + t], s0);
    uv[t] = s0 + s1;
  }
  __syncthreads();

  for (int j = 0; j < k; ++j) {
    {  // X' u_j: thread = (column a, time split sp), eight loads in flight
      const int a = tid % astr, sp = tid / astr;
      if (sp < nsplit) {
        const int chunk = (n + nsplit - 1) / nsplit;
        const int t0 = sp * chunk, t1 = min(n, t0 + chunk);
        double s[8];
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) s[q8] = 0.0;
        if (a < d) {
          const size_t ldX = D.ldX;
          const double* xp = Xw + (size_t)t0 * ldX + a;
          int t = t0;
          for (; t + 7 < t1; t += 8, xp += 8 * ldX) {
            double xv[8];
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) xv[q8] = xp[q8 * ldX];
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) s[q8] = fma(xv[q8], uv[t + q8], s[q8]);
          }
          for (; t < t1; ++t, xp += ldX) s[0] = fma(xp[0], uv[t], s[0]);
        }
        part[sp * astr + a] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
      }
    }
    __syncthreads();
    for (int a = tid; a < d; a += nt) {
      double acc = prec_u[j * d + a] * (P.prior_mean[j * d + a] - Ac[(size_t)a * ldM + j]);
      for (int sp = 0; sp < nsplit; ++sp) acc += part[sp * astr + a];
      yv[a] = acc;
      zv[a] = rng.normal_elem(ST_COEF_Z, (uint32_t)(j * d + a));
    }
    if (FRES) {
      cp_async_wait<0>();
      precision_solve_smem(big, d, yv, zv);  // starts and ends with __syncthreads
      if (j + 1 < k) prefetch_factor(j + 1);
    } else {
      precision_solve_cta(Fg + (size_t)j * D.ldS, d, yv, zv, blk);
    }
    for (int a = tid; a < d; a += nt) {
      const double val = Ac[(size_t)a * ldM + j] + yv[a];
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
    if (j + 1 < k) {
      {  // q = X delta_j: thread = (time t, column split sa), partial sums through shared memory
        const int t = tid % tstr, sa = tid / tstr;
        if (sa < nsa) {
          const int chunk = (d + nsa - 1) / nsa;
          const int a0 = sa * chunk, a1 = min(d, a0 + chunk);
          double s[8];
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) s[q8] = 0.0;
          if (t < n) {
            const size_t ldT = D.ldT;
            const double* xp = XTw + (size_t)a0 * ldT + t;
            int a = a0;
            for (; a + 7 < a1; a += 8, xp += 8 * ldT) {
              double xv[8];
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) xv[q8] = xp[q8 * ldT];
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) s[q8] = fma(xv[q8], yv[a + q8], s[q8]);
            }
            for (; a < a1; ++a, xp += ldT) s[0] = fma(xp[0], yv[a], s[0]);
          }
          part[sa * tstr + t] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        }
      }
      __syncthreads();
      for (int t = tid; t < n; t += nt) {  // E -= L_.j q, then u_{j+1}
        double qt = 0.0;
        for (int sa = 0; sa < nsa; ++sa) qt += part[sa * tstr + t];
        double un0 = 0.0, un1 = 0.0;
        int i = j + 1;
        for (; i + 3 < k; i += 4) {
          double e[4], dv[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            e[q4] = ET[(size_t)(i + q4) * npad + t];
            dv[q4] = DT[(size_t)(i + q4) * npad + t];
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            e[q4] = fma(-Lsm[(i + q4) * k + j], qt, e[q4]);
            ET[(size_t)(i + q4) * npad + t] = e[q4];
          }
          un0 = fma(Lsm[i * k + j + 1] * e[0], dv[0], un0);
          un1 = fma(Lsm[(i + 1) * k + j + 1] * e[1], dv[1], un1);
          un0 = fma(Lsm[(i + 2) * k + j + 1] * e[2], dv[2], un0);
          un1 = fma(Lsm[(i + 3) * k + j + 1] * e[3], dv[3], un1);
        }
        for (; i < k; ++i) {
          const size_t at = (size_t)i * npad + t;
          const double e = fma(-Lsm[i * k + j], qt, ET[at]);
          ET[at] = e;
          un0 = fma(Lsm[i * k + j + 1] * e, DT[at], un0);
        }
        uv[t] = un0 + un1;
      }
    }
    __syncthreads();
  }
}
#endif  // BV_NC == 8

}  // namespace

cudaError_t BV_CAT(launch_seq_big_nc, BV_NC)(const Dims& D, const DevPtrs& P, const double* Pfac, double* hscr, size_t smem, cudaStream_t st,
                                             const double* Rres) {
  if (D.sv) {
    cudaFuncSetAttribute(coef_seq_big_kernel<true, BV_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_seq_big_kernel<true, BV_NC><<<D.U, CB_THREADS, smem, st>>>(D, P, Pfac, hscr, nullptr);
  } else {
    cudaFuncSetAttribute(coef_seq_big_kernel<false, BV_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_seq_big_kernel<false, BV_NC><<<D.U, CB_THREADS, smem, st>>>(D, P, Pfac, hscr, Rres);
  }
  return cudaGetLastError();
}


#if BV_NC == 8
static bool seq_time_resident(const Dims& D, size_t smem_limit) {
  const int npad = (D.nmax + 3) & ~3;
  return td_layout(D.k, D.d, npad, D.ldS, true).total * sizeof(double) <= smem_limit;
}
size_t seq_time_smem(const Dims& D) {  // smallest configuration (factor read from global memory)
  return td_layout(D.k, D.d, (D.nmax + 3) & ~3, D.ldS, false).total * sizeof(double);
}
size_t seq_time_scratch_doubles(const Dims& D) { return (size_t)D.U * 2 * D.k * ((D.nmax + 3) & ~3); }
cudaError_t launch_seq_time(const Dims& D, const DevPtrs& P, const double* Pfac, double* escr, size_t smem_limit, cudaStream_t st) {
  const int npad = (D.nmax + 3) & ~3;
  if (seq_time_resident(D, smem_limit)) {
    const size_t smem = td_layout(D.k, D.d, npad, D.ldS, true).total * sizeof(double);
    cudaFuncSetAttribute(coef_seq_time_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_seq_time_kernel<true><<<D.U, TD_THREADS, smem, st>>>(D, P, Pfac, escr, npad);
    return cudaGetLastError();
  }
  const size_t smem = seq_time_smem(D);
  if (smem > smem_limit) return cudaErrorInvalidConfiguration;
  cudaFuncSetAttribute(coef_seq_time_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  coef_seq_time_kernel<false><<<D.U, TD_THREADS, smem, st>>>(D, P, Pfac, escr, npad);
  return cudaGetLastError();
}
#endif

}  // namespace bvhar
