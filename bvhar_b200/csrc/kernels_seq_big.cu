// kernels_seq_big.cu - the sequential part of the coefficient stage for large designs (see kernels_chol_big.cu);
// compiled once per BV_NC = 8 | 16 (ceil(d / 32) class).
#include "block_linalg.cuh"
#include "engine.cuh"
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
__device__ void precision_solve_cta(const double* __restrict__ Lg, int d, double* y, const double* z, double* blk /* 32 x 33 */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int nb = (d + 31) / 32;
  // forward: L y = r
  for (int jb = 0; jb < nb; ++jb) {
    const int c0 = 32 * jb, nc = min(32, d - c0);
    __syncthreads();
    for (int e = tid; e < 32 * 32; e += nt) {  // diagonal block rows into shared memory
      const int rr = e >> 5, cc = e & 31;
      blk[rr * 33 + cc] = (rr < nc && cc <= rr) ? Lg[rowoff(c0 + rr) + c0 + cc] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < nc ? y[c0 + lane] : 0.0;
      const double inv = lane < nc ? 1.0 / blk[lane * 33 + lane] : 0.0;
      v *= inv;
      for (int c = 0; c < nc; ++c) {
        const double yc = __shfl_sync(0xffffffffu, v, c);
        if (lane > c && lane < nc) v = fma(-blk[lane * 33 + c] * inv, yc, v);
      }
      if (lane < nc) y[c0 + lane] = v;
    }
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
  __syncthreads();
  for (int a = tid; a < d; a += nt) y[a] += z[a];
  // backward: L' x = y + z
  for (int jb = nb - 1; jb >= 0; --jb) {
    const int c0 = 32 * jb, nc = min(32, d - c0);
    __syncthreads();
    for (int e = tid; e < 32 * 32; e += nt) {
      const int rr = e >> 5, cc = e & 31;
      blk[rr * 33 + cc] = (rr < nc && cc <= rr) ? Lg[rowoff(c0 + rr) + c0 + cc] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < nc ? y[c0 + lane] : 0.0;
      const double inv = lane < nc ? 1.0 / blk[lane * 33 + lane] : 0.0;
      v *= inv;
      for (int c = nc - 1; c >= 0; --c) {
        const double xc = __shfl_sync(0xffffffffu, v, c);
        if (lane < c) v = fma(-blk[c * 33 + lane] * inv, xc, v);
      }
      if (lane < nc) y[c0 + lane] = v;
    }
    __syncthreads();
    for (int b = tid; b < c0; b += nt) {  // columns to the left: y_b -= sum_a L[c0 + a][b] x_a   (coalesced in b)
      double s0 = 0.0, s1 = 0.0;
      for (int a = 0; a + 1 < nc; a += 2) {
        s0 = fma(Lg[rowoff(c0 + a) + b], y[c0 + a], s0);
        s1 = fma(Lg[rowoff(c0 + a + 1) + b], y[c0 + a + 1], s1);
      }
      if (nc & 1) s0 = fma(Lg[rowoff(c0 + nc - 1) + b], y[c0 + nc - 1], s0);
      y[b] -= s0 + s1;
    }
  }
  __syncthreads();
}

// ---- B (large d): the sequential-in-j part, one CTA per unit ----------------------------------------------------
template <bool SV, int NC>
__global__ void __launch_bounds__(CB_THREADS) coef_seq_big_kernel(const Dims D, const DevPtrs P, const double* __restrict__ Pfac,
                                                                   double* __restrict__ hscr) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int k = D.k, d = D.d, ldM = D.ldM;
  double* Lsm = sm;                          // [k][k]
  double* yv = Lsm + (size_t)k * k;          // [d] rhs -> delta
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

  for (int e = tid; e < k * k; e += blockDim.x) Lsm[e] = P.L[(size_t)u * k * k + e];
  if (!SV) for (int i = tid; i < k; i += blockDim.x) idg[i] = 1.0 / P.diag[(size_t)u * k + i];
  __syncthreads();
  for (int i = warp; i < k; i += nw) {  // h_i = C_i - S_i v_i, v_i = A L_i'
    double* v = wscr + (size_t)warp * d;
    for (int a = lane; a < d; a += 32) {
      double sacc = 0.0;
      for (int m = 0; m <= i; ++m) sacc = fma(Lsm[i * k + m], Ac[(size_t)a * ldM + m], sacc);
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
          for (int m = 0; m <= i; ++m) sacc = fma(xty[m], Lsm[i * k + m], sacc);
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
      for (int i = j; i < k; ++i) acc = fma(Lsm[i * k + j], h[(size_t)i * d + a], acc);
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
          const double lij = Lsm[i * k + j];
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) h[(size_t)i * d + b] = fma(-lij, out[sidx], h[(size_t)i * d + b]);
          }
        }
      } else {
        // X'X delta once, rows split over the warps: warp q takes the 32-row blocks q, q + nw, ...
        for (int a = tid; a < d; a += blockDim.x) tv[a] = 0.0;
        __syncthreads();
        {
          double out[NC];
          symv_packed_warp<NC>(X2, yv, d, out, warp, nw);
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d && out[sidx] != 0.0) atomicAdd(&tv[b], out[sidx]);
          }
        }
        __syncthreads();
        for (int e = tid; e < (k - j - 1) * d; e += blockDim.x) {
          const int i = j + 1 + e / d, a = e % d;
          h[(size_t)i * d + a] = fma(-Lsm[i * k + j] * idg[i], tv[a], h[(size_t)i * d + a]);
        }
      }
    }
    __threadfence_block();
    __syncthreads();
  }
}

}  // namespace

cudaError_t BV_CAT(launch_seq_big_nc, BV_NC)(const Dims& D, const DevPtrs& P, const double* Pfac, double* hscr, size_t smem, cudaStream_t st) {
  if (D.sv) {
    cudaFuncSetAttribute(coef_seq_big_kernel<true, BV_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_seq_big_kernel<true, BV_NC><<<D.U, CB_THREADS, smem, st>>>(D, P, Pfac, hscr);
  } else {
    cudaFuncSetAttribute(coef_seq_big_kernel<false, BV_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    coef_seq_big_kernel<false, BV_NC><<<D.U, CB_THREADS, smem, st>>>(D, P, Pfac, hscr);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
