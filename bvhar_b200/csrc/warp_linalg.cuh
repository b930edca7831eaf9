// warp_linalg.cuh — warp-level dense kernels on packed-lower (row-major packed) symmetric matrices:
// the batched small-system building blocks behind draw_coef (bayes/misc/draw.h:372-383) for thousands
// of (unit, equation) systems at once.
//
//   chol_blocked_warp<NB>      in-place Cholesky of a matrix in shared memory, d <= 32 NB.  Right-looking over
//                              32 x 32 blocks; a block lives in REGISTERS, one row per lane, so the inner
//                              updates are FMA + shuffle with static register indices (no address math,
//                              no predication): ~10k warp instructions for d = 61 instead of ~66k for the
//                              element-wise shared-memory version it replaces (ncu, profiles/).
//   precision_solve_lanes<NC>  delta = L^-T (L^-1 r + z): lanes own rows, one shuffle per column.
//   symv_packed_warp<NC>       y = S x with S streamed row by row from global / L2 memory (coalesced).
#pragma once
#include <cuda_runtime.h>

namespace bvhar {

__device__ __forceinline__ double wl_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#define wsum wl_wsum

template <int NB>
__device__ __forceinline__ void chol_blocked_warp(double* Pk, int d, int* bad) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double dg[32], wk[32];
  for (int jb = 0; jb < NB; ++jb) {
    if (32 * jb >= d) break;
    const int a = 32 * jb + lane;
    const bool va = a < d;
    const int ra = a * (a + 1) / 2 + 32 * jb;
    // ---- diagonal block: load, factor in registers, store
#pragma unroll
    for (int c = 0; c < 32; ++c) dg[c] = (va && c <= lane) ? Pk[ra + c] : (c == lane ? 1.0 : 0.0);
    double dinv = 1.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const double pc = __shfl_sync(FULL, dg[c], c);
      const double piv = sqrt(pc), inv = 1.0 / piv;
      if (lane == c && !(pc > 0.0)) *bad = 1;
      const double lc = lane > c ? dg[c] * inv : 0.0;
      if (lane == c) { dg[c] = piv; dinv = inv; }
      else dg[c] = lc;
#pragma unroll
      for (int b = c + 1; b < 32; ++b) {
        const double lbc = __shfl_sync(FULL, lc, b);
        dg[b] = fma(-lc, lbc, dg[b]);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (va && c <= lane) Pk[ra + c] = dg[c];
    // ---- panel: L[ib][jb] = A[ib][jb] L[jb][jb]^-T, one block row at a time
    for (int ib = jb + 1; ib < NB; ++ib) {
      if (32 * ib >= d) break;
      const int a2 = 32 * ib + lane;
      const bool v2 = a2 < d;
      const int r2 = a2 * (a2 + 1) / 2 + 32 * jb;
#pragma unroll
      for (int c = 0; c < 32; ++c) wk[c] = v2 ? Pk[r2 + c] : 0.0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double xc = wk[c] * __shfl_sync(FULL, dinv, c);
        wk[c] = xc;
#pragma unroll
        for (int b = c + 1; b < 32; ++b) {
          const double lbc = __shfl_sync(FULL, dg[c], b);  // L[jb][jb](b, c): lane b, register c
          wk[b] = fma(-xc, lbc, wk[b]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (v2) Pk[r2 + c] = wk[c];
    }
    __syncwarp();
    // ---- trailing update: A[ib][kb] -= L[ib][jb] L[kb][jb]^T for ib >= kb > jb
    for (int kb = jb + 1; kb < NB; ++kb) {
      if (32 * kb >= d) break;
      for (int ib = kb; ib < NB; ++ib) {
        if (32 * ib >= d) break;
        const int a2 = 32 * ib + lane;
        const bool v2 = a2 < d;
        const int rl = a2 * (a2 + 1) / 2 + 32 * jb;  // row a2 of L[ib][jb]
        const int rw = a2 * (a2 + 1) / 2 + 32 * kb;  // row a2 of A[ib][kb]
#pragma unroll
        for (int c = 0; c < 32; ++c) dg[c] = v2 ? Pk[rl + c] : 0.0;
#pragma unroll
        for (int b = 0; b < 32; ++b) wk[b] = (v2 && 32 * kb + b <= a2) ? Pk[rw + b] : 0.0;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
          const int ab = 32 * kb + b;
          if (ab < d) {  // warp-uniform
            const double* lk = Pk + ab * (ab + 1) / 2 + 32 * jb;  // row ab of L[kb][jb], same address in every lane
            double acc = wk[b];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc = fma(-dg[c], lk[c], acc);
            wk[b] = acc;
          }
        }
#pragma unroll
        for (int b = 0; b < 32; ++b)
          if (v2 && 32 * kb + b <= a2) Pk[rw + b] = wk[b];
      }
      __syncwarp();
    }
  }
  __syncwarp();
}

// out (lane-distributed: element b = lane + 32 c in out[c]) = S x, S packed lower (row-major packed) in
// global memory, x in shared memory.  NC = ceil(d / 32).
template <int NC>
__device__ __forceinline__ void symv_packed_warp(const double* __restrict__ S, const double* x, int d, double (&out)[NC]) {
  const int lane = threadIdx.x & 31;
  double xb[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int b = lane + 32 * c;
    xb[c] = b < d ? x[b] : 0.0;
    out[c] = 0.0;
  }
  constexpr int UR = 4;
  for (int a0 = 0; a0 < d; a0 += UR) {
    double s[UR][NC];
#pragma unroll
    for (int r = 0; r < UR; ++r) {
      const int a = a0 + r;
      const int ra = a * (a + 1) / 2;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int b = lane + 32 * c;
        s[r][c] = (a < d && b <= a) ? S[ra + b] : 0.0;
      }
    }
#pragma unroll
    for (int r = 0; r < UR; ++r) {
      const int a = a0 + r;
      if (a >= d) break;
      const double xa = x[a];
      double part = 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int b = lane + 32 * c;
        part = fma(s[r][c], xb[c], part);
        if (b < a) out[c] = fma(s[r][c], xa, out[c]);
      }
      part = wsum(part);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if ((a >> 5) == c && (a & 31) == lane) out[c] += part;
    }
  }
}

// Warp Cholesky of a packed-lower matrix in shared memory (right-looking; lanes own columns of the
// trailing update, their scaled column entries stay in registers).
template <int NC>
__device__ __forceinline__ void chol_warp(double* Pk, int d, int* bad) {
  const int lane = threadIdx.x & 31;
  for (int c = 0; c < d; ++c) {
    const int cc = c * (c + 1) / 2 + c;
    const double p = Pk[cc];
    const double piv = sqrt(p), inv = 1.0 / piv;
    double la[NC];
#pragma unroll
    for (int s = 0; s < NC; ++s) {
      const int a = lane + 32 * s;
      la[s] = 0.0;
      if (a > c && a < d) {
        la[s] = Pk[a * (a + 1) / 2 + c] * inv;
        Pk[a * (a + 1) / 2 + c] = la[s];
      }
    }
    __syncwarp();
    if (lane == 0) {
      Pk[cc] = piv;
      if (!(p > 0.0)) *bad = 1;
    }
    constexpr int UR = 4;  // rows are independent: batch the loads of UR rows ahead of their stores
    for (int a0 = c + 1; a0 < d; a0 += UR) {
      double lac[UR], cur[UR][NC];
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        const int a = a0 + r;
        const int ra = a * (a + 1) / 2;
        lac[r] = a < d ? Pk[ra + c] : 0.0;
#pragma unroll
        for (int s = 0; s < NC; ++s) {
          const int b = lane + 32 * s;
          cur[r][s] = (a < d && b > c && b <= a) ? Pk[ra + b] : 0.0;
        }
      }
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        const int a = a0 + r;
        const int ra = a * (a + 1) / 2;
#pragma unroll
        for (int s = 0; s < NC; ++s) {
          const int b = lane + 32 * s;
          if (a < d && b > c && b <= a) Pk[ra + b] = fma(-lac[r], la[s], cur[r][s]);
        }
      }
    }
    __syncwarp();
  }
}

// delta = L^-T (L^-1 r + z) by one warp; L packed lower in shared memory, r / z / out in shared memory.
template <int NC>
__device__ __forceinline__ void precision_solve_lanes(const double* Lp, int d, const double* r, const double* z, double* out) {
  const int lane = threadIdx.x & 31;
  double y[NC], inv[NC];
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    const int a = lane + 32 * s;
    y[s] = a < d ? r[a] : 0.0;
    inv[s] = a < d ? 1.0 / Lp[a * (a + 1) / 2 + a] : 0.0;
  }
  for (int c = 0; c < d; ++c) {  // forward: lanes own rows
    double yc = 0.0;
#pragma unroll
    for (int s = 0; s < NC; ++s)
      if ((c >> 5) == s) yc = y[s] * inv[s];
    yc = __shfl_sync(0xffffffffu, yc, c & 31);
#pragma unroll
    for (int s = 0; s < NC; ++s) {
      const int a = lane + 32 * s;
      if (a == c) y[s] = yc;
      else if (a > c && a < d) y[s] = fma(-Lp[a * (a + 1) / 2 + c], yc, y[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    const int a = lane + 32 * s;
    if (a < d) y[s] += z[a];
  }
  for (int c = d - 1; c >= 0; --c) {  // backward with L^T: lanes own entries b, row c of L is contiguous
    double xc = 0.0;
#pragma unroll
    for (int s = 0; s < NC; ++s)
      if ((c >> 5) == s) xc = y[s] * inv[s];
    xc = __shfl_sync(0xffffffffu, xc, c & 31);
    const int rc = c * (c + 1) / 2;
#pragma unroll
    for (int s = 0; s < NC; ++s) {
      const int b = lane + 32 * s;
      if (b == c) y[s] = xc;
      else if (b < c) y[s] = fma(-Lp[rc + b], xc, y[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    const int a = lane + 32 * s;
    if (a < d) out[a] = y[s];
  }
}


#undef wsum
}  // namespace bvhar
