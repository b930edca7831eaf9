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

// Cholesky of one 32 x 32 block held in registers, one row per lane (dg[c] = A[lane][c], c <= lane; rows >= nv must be
// passed as identity rows), with the column broadcasts through SHARED MEMORY instead of warp shuffles: after the pivot
// step every lane drops its entry of column c into a 32-double buffer (double-buffered by the parity of c, one
// __syncwarp per step) and the rank-1 update reads L[b][c], b > c, from there - one broadcast 8-byte load with an
// immediate offset per value instead of two 32-bit shuffles.  The shuffle version issues 2 x 496 SHFL per block and is
// bound by the shuffle pipe; this one is bound by the pivot chain.  colbuf: 64 doubles of shared memory owned by the
// warp.  On exit dg holds L; returns 1 / L[lane][lane].
__device__ __forceinline__ double chol32_smem(double (&dg)[32], double* colbuf, int* bad) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double dinv = 1.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const double pc = __shfl_sync(FULL, dg[c], c);
    const double inv = rsqrt(pc), piv = pc * inv;
    if (lane == c && !(pc > 0.0)) *bad = 1;
    const double lc = lane > c ? dg[c] * inv : 0.0;
    if (lane == c) { dg[c] = piv; dinv = inv; }
    else dg[c] = lc;
    double* cb = colbuf + 32 * (c & 1);
    cb[lane] = lc;
    __syncwarp();
#pragma unroll
    for (int b = c + 1; b < 32; ++b) dg[b] = fma(-lc, cb[b], dg[b]);
  }
  __syncwarp();
  return dinv;
}

// The same factorisation for a warp that runs ALONE (the CTA-per-matrix Cholesky of large designs: the other warps wait
// for this block): there the 32-step pivot chain is the critical path, and the next pivot does not need the column
// broadcast - lane c+1 updates its own diagonal entry with its own multiplier (bit-identical to what the broadcast
// update writes a moment later) and shuffles it out before the shared-memory round trip, so a step of the chain is
// shuffle + rsqrt + multiply + FMA instead of that plus store, warp barrier and load.
__device__ __forceinline__ double chol32_smem_chain(double (&dg)[32], double* colbuf, int* bad) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double dinv = 1.0;
  double pc = __shfl_sync(FULL, dg[0], 0);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const double inv = rsqrt(pc), piv = pc * inv;
    if (lane == c && !(pc > 0.0)) *bad = 1;
    const double lc = lane > c ? dg[c] * inv : 0.0;
    if (lane == c) { dg[c] = piv; dinv = inv; }
    else dg[c] = lc;
    if (c < 31) pc = __shfl_sync(FULL, fma(-lc, lc, dg[c + 1]), c + 1);  // next pivot, ahead of the broadcast
    double* cb = colbuf + 32 * (c & 1);
    cb[lane] = lc;
    __syncwarp();
#pragma unroll
    for (int b = c + 1; b < 32; ++b) dg[b] = fma(-lc, cb[b], dg[b]);
  }
  __syncwarp();
  return dinv;
}

// Column `c` (one thread per column) of the inverse of a 32 x 32 lower-triangular block whose rows are read through
// `Lrow(r)` (pointer to L[r][0], the same address for every thread => shared-memory broadcasts) and whose reciprocal
// diagonal is dinv[r]: forward substitution L x = e_c, 496 predicated FMAs per thread with two partial sums.  Rows >= nv
// are identity rows.
template <typename RowFn>
__device__ __forceinline__ void inv32_column(RowFn Lrow, const double* dinv, int c, int nv, double (&x)[32]) {
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const double* lr = Lrow(r);
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m + 1 < r; m += 2) {
      s0 = fma(-lr[m], x[m], s0);
      s1 = fma(-lr[m + 1], x[m + 1], s1);
    }
    if (r & 1) s0 = fma(-lr[r - 1], x[r - 1], s0);
    x[r] = (r < c) ? 0.0 : (r < nv ? (s0 + s1) * dinv[r] : ((r == c) ? 1.0 : 0.0));
  }
}

template <int NB>
__device__ __forceinline__ void chol_blocked_warp(double* Pk, int d, int* bad, double* colbuf /* 64 doubles of shared memory per warp */) {
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
    const double dinv = chol32_smem(dg, colbuf, bad);
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (va && c <= lane) Pk[ra + c] = dg[c];
    // ---- panel: L[ib][jb] = A[ib][jb] L[jb][jb]^-T, one block row at a time; column c of the diagonal block (lane b
    // holds L[jb][jb](b, c) in dg[c]) is broadcast through colbuf
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
        double* cbf = colbuf + 32 * (c & 1);
        cbf[lane] = dg[c];
        __syncwarp();
#pragma unroll
        for (int b = c + 1; b < 32; ++b) wk[b] = fma(-xc, cbf[b], wk[b]);
      }
      __syncwarp();
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
// Rows are taken 16 at a time with ALL their loads in flight together (16 (rb + 1) independent 8-byte loads per
// lane: the product is latency-bound, S comes from L2 / HBM).  Lane b accumulates the row partials S_ab x_b in
// registers (part[r], static indices) and its own column contribution S_ab x_a; one butterfly reduce-scatter per
// 16 rows (16 shuffles instead of 5 per row) leaves the total of row 16 g + (l & 15) on lanes l and l ^ 16.
// With (first, step) a warp takes only the 32-row blocks rb = first, first + step, ... (its `out` is then a partial
// result to be summed over the cooperating warps).
template <int NC>
__device__ __forceinline__ void symv_packed_warp(const double* __restrict__ S, const double* x, int d, double (&out)[NC],
                                                 int first = 0, int step = 1) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double xb[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int b = lane + 32 * c;
    xb[c] = b < d ? x[b] : 0.0;
    out[c] = 0.0;
  }
#pragma unroll
  for (int rb = 0; rb < NC; ++rb) {
    if (32 * rb >= d) break;
    if (step > 1 && (rb % step) != first) continue;  // warp-uniform
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int a0 = 32 * rb + 16 * hb;
      if (a0 >= d) break;
      double s[16][NC];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int a = a0 + r;
        const int ra = a * (a + 1) / 2;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int b = lane + 32 * c;
          s[r][c] = (c <= rb && a < d && b <= a) ? S[ra + b] : 0.0;
        }
      }
      double part[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int a = a0 + r;
        const double xa = a < d ? x[a] : 0.0;
        double p = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c > rb) continue;
          const int b = lane + 32 * c;
          p = fma(s[r][c], xb[c], p);
          out[c] = fma(s[r][c], b < a ? xa : 0.0, out[c]);
        }
        part[r] = p;
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const double send = up ? part[i] : part[i + off];
          const double keep = up ? part[i + off] : part[i];
          part[i] = keep + __shfl_xor_sync(FULL, send, off);
        }
      }
      const double tot = part[0] + __shfl_xor_sync(FULL, part[0], 16);  // row a0 + (lane & 15)
      if ((lane >> 4) == hb) out[rb] += tot;
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
    const double inv = rsqrt(p), piv = p * inv;  // one MUFU + Newton instead of sqrt followed by a division
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


// delta = L^-T (L^-1 r + z) by one warp with the 32 x 32 diagonal blocks of L held in REGISTERS (lane = row
// for the forward solve, lane = column for the backward solve), rows pre-scaled by 1 / L_aa so that the
// dependent chain of each of the 2 d substitution steps is one shuffle + one FMA; the off-diagonal block
// updates are independent dot products.  `out` (shared memory, d doubles) doubles as the broadcast buffer.
template <int NC>
__device__ __forceinline__ void precision_solve_blocked(const double* Lp, int d, const double* r, const double* z, double* out) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double y[NC], inv[NC];
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    const int a = lane + 32 * s;
    inv[s] = a < d ? 1.0 / Lp[a * (a + 1) / 2 + a] : 0.0;
    y[s] = a < d ? r[a] * inv[s] : 0.0;  // scaled unknowns: y_a / L_aa
  }
  // ---- forward: L y = r
#pragma unroll
  for (int jb = 0; jb < NC; ++jb) {
    if (32 * jb >= d) break;
    const int a = 32 * jb + lane;
    const int ra = a * (a + 1) / 2 + 32 * jb;
    double dg[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) dg[c] = (a < d && c < lane) ? Lp[ra + c] * inv[jb] : 0.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const double yc = __shfl_sync(FULL, y[jb], c);
      y[jb] = fma(-dg[c], yc, y[jb]);  // dg[c] = 0 for lanes <= c: solved entries stay put
    }
    if (a < d) out[a] = y[jb];
    __syncwarp();
#pragma unroll
    for (int ib = jb + 1; ib < NC; ++ib) {
      const int a2 = 32 * ib + lane;
      if (a2 < d) {
        const double* row = Lp + a2 * (a2 + 1) / 2 + 32 * jb;
        double acc = 0.0;
#pragma unroll 8
        for (int c = 0; c < 32; ++c) acc = fma(row[c], out[32 * jb + c], acc);
        y[ib] = fma(-acc, inv[ib], y[ib]);
      }
    }
    __syncwarp();
  }
  // ---- backward: L' x = y + z.  y holds (L^-1 r)_a exactly; pre-scale the right-hand side by 1 / L_cc
#pragma unroll
  for (int s = 0; s < NC; ++s) {
    const int a = lane + 32 * s;
    y[s] = a < d ? (y[s] + z[a]) * inv[s] : 0.0;
  }
#pragma unroll
  for (int jb = NC - 1; jb >= 0; --jb) {
    if (32 * jb >= d) continue;
    const int b = 32 * jb + lane;
    double dgT[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int a = 32 * jb + c;
      dgT[c] = (a < d && c > lane) ? Lp[a * (a + 1) / 2 + b] * inv[jb] : 0.0;
    }
#pragma unroll
    for (int c = 31; c >= 0; --c) {
      const double xc = __shfl_sync(FULL, y[jb], c);
      y[jb] = fma(-dgT[c], xc, y[jb]);
    }
    if (b < d) out[b] = y[jb];
    __syncwarp();
#pragma unroll
    for (int ib = 0; ib < jb; ++ib) {
      const int b2 = 32 * ib + lane;  // always < d here
      double acc = 0.0;
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        const int a = 32 * jb + c;
        if (a < d) acc = fma(Lp[a * (a + 1) / 2 + b2], out[a], acc);
      }
      y[ib] = fma(-acc, inv[ib], y[ib]);
    }
    __syncwarp();
  }
}

#undef wsum
}  // namespace bvhar
