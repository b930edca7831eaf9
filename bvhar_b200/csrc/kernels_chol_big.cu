// kernels_chol_big.cu (+ kernels_seq_big.cu) — stage 4 (updateCoef, triangular.h:281-316 + draw_coef draw.h:372-383) for LARGE designs
// (dim_design > 96: BASELINE configs C4, d = 201, and C5, d = 301), where one packed precision matrix is too big
// for the one-warp-per-matrix / everything-in-shared-memory kernels of kernels_chol.cu / kernels_coef.cu:
//
//   coef_chol_cta_kernel   chol(diag(prec_j) + P_j) for every (unit, equation), ONE CTA PER MATRIX, in place in
//                          global memory (the 148 x <= 2 matrices in flight stay L2-resident).  Left-looking
//                          blocked factorisation, 32-column panels: the panel update
//                              A[:, jb] -= L[:, <jb] L[jb, <jb]'
//                          holds ~all of the d^3/3 flops and runs on the fp64 tensor pipe (DMMA m8n8k4, A fragments
//                          straight from L2, B fragments from a shared-memory tile); the 32 x 32 diagonal block is
//                          factored in registers by one warp (warp_linalg.cuh) and the rows below it are solved
//                          one thread per row.
//   invert_diag_kernel     replaces every 32 x 32 DIAGONAL BLOCK of the factors by its INVERSE (lower triangular),
//                          one warp per block, one thread per column (forward substitution L x = e_c from a
//                          shared-memory copy of the block): the blocked triangular solves of kernels_seq_big.cu are
//                          then matrix-vector products instead of 32-step dependent substitutions.  This is the
//                          storage convention of Pfac on the large-design path.
//   coef_seq_big_kernel    the part that is sequential in j (one CTA per unit) with the factor read from
//                          global / L2 by a blocked two-sided triangular solve that uses all warps, and the
//                          k x d work vectors h_i = C_i - S_i v_i in a global scratch.
#include <algorithm>
#include <cstdlib>

#include "block_linalg.cuh"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "warp_linalg.cuh"

namespace bvhar {

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_PB = 36;  // pitch of the 32 x 32 B tile: 36 = 4 (mod 16) => conflict-free DMMA fragment loads

__device__ __forceinline__ size_t rowoff(int a) { return (size_t)a * (a + 1) / 2; }

// ---- A2 (large d): one CTA per matrix --------------------------------------------------------------------
// MTW = m-tiles (8 rows) per warp: ceil(ceil(d / 8) / 8).
template <int MTW>
__global__ void __launch_bounds__(CB_THREADS, (MTW <= 4 ? 2 : 1)) coef_chol_cta_kernel(const Dims D, const DevPtrs P, double* Pfac) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const long long job = blockIdx.x;
  const int u = (int)(job / D.k), j = (int)(job % D.k), w = u / D.C;
  const int k = D.k, d = D.d;
  const bool SV = D.sv != 0;
  double* pan = sm;                       // [d][33] current panel, rows c0 .. d-1
  double* bt = pan + (size_t)d * 33;      // [32][CB_PB] rows of the diagonal block, 32 previous columns at a time
  double* red = bt + 32 * CB_PB;          // [64] scratch
  double* pan2 = red + 64;                // [d - 32][33] the next panel, prefetched
  __shared__ int bad;
  __shared__ double colbuf[64];  // column broadcast buffer of the diagonal-block factorisation (warp_linalg.cuh)
  if (tid == 0) bad = 0;
  double* Lg = Pfac + ((size_t)u * k + j) * D.ldS;  // in/out (SV: holds P_j on entry)
  const double* src = Lg;
  double scale = 1.0;
  if (!SV) {  // P_j = (sum_{i>=j} L_ij^2 / d_i) X'X
    double ws = 0.0;
    for (int i = j + tid; i < k; i += CB_THREADS) {
      const double l = P.L[((size_t)u * k + i) * k + j];
      ws += l * l / P.diag[(size_t)u * k + i];
    }
    scale = block_sum(ws, red);
    src = P.XtX + (size_t)w * D.ldS;
  }
  const double* pr = P.prec + (size_t)u * k * d + (size_t)j * d;

  for (int c0 = 0; c0 < d; c0 += 32) {
    const int m = d - c0;            // rows of the panel
    const int nc = min(32, m);       // columns of the panel
    __syncthreads();
    // 1. the panel (rows r = c0 + rr, columns c0 .. min(c0 + 31, r)) with the prior precision on the diagonal: read from
    //    global memory for the first panel, afterwards taken from the copy that warps 1..7 prefetched into pan2 while
    //    warp 0 factored the previous diagonal block (step 3)
    if (c0 == 0) {
      for (int e = tid; e < m * 32; e += CB_THREADS) {
        const int rr = e >> 5, cc = e & 31;
        const int r = c0 + rr, c = c0 + cc;
        double v = 0.0;
        if (c <= r && cc < nc) {
          v = scale * src[rowoff(r) + c];
          if (c == r) v += pr[r];
        }
        pan[rr * 33 + cc] = v;
      }
    } else {
      for (int e = tid; e < m * 33; e += CB_THREADS) pan[e] = pan2[e];
    }
    // 2. panel -= L[rows, 0:c0] * L[c0:c0+32, 0:c0]'   (DMMA; K = c0 in chunks of 32)
    if (c0 > 0) {
      double acc[MTW][4][2];
#pragma unroll
      for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
      const int mtiles = (m + 7) / 8;
      for (int kk0 = 0; kk0 < c0; kk0 += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * 32; e += CB_THREADS) {  // bt[n][kk] = L[c0 + n][kk0 + kk]
          const int nn = e >> 5, kk = e & 31;
          bt[nn * CB_PB + kk] = (c0 + nn < d) ? Lg[rowoff(c0 + nn) + kk0 + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MTW; ++i) {
          const int mt = warp + 8 * i;
          if (mt >= mtiles) continue;  // warp-uniform
          const int r = c0 + 8 * mt + fr;
          const double* Lr = Lg + rowoff(min(r, d - 1)) + kk0 + fc;
          double af[8];
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) af[ks] = r < d ? Lr[4 * ks] : 0.0;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[i][nt][0], acc[i][nt][1], af[ks], bt[(8 * nt + fr) * CB_PB + 4 * ks + fc]);
        }
      }
#pragma unroll
      for (int i = 0; i < MTW; ++i) {  // fragments: row 8 mt + fr, columns 8 nt + 2 fc + {0, 1}
        const int mt = warp + 8 * i;
        if (mt >= mtiles) continue;
        const int rr = 8 * mt + fr;
        if (rr < m) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            pan[rr * 33 + 8 * nt + 2 * fc] -= acc[i][nt][0];
            pan[rr * 33 + 8 * nt + 2 * fc + 1] -= acc[i][nt][1];
          }
        }
      }
    }
    __syncthreads();
    // 3. factor the diagonal block: one warp, the block in registers (one row per lane)
    if (warp == 0) {
      double dg[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) dg[c] = (lane < nc && c <= lane) ? pan[lane * 33 + c] : (c == lane ? 1.0 : 0.0);
      const double dinv = chol32_smem_chain(dg, colbuf, &bad);
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (lane < nc && c <= lane) pan[lane * 33 + c] = dg[c];
      red[lane] = dinv;  // 1 / L_cc for the row solves
    } else if (c0 + 32 < d) {  // the other warps: fetch the next panel's source values (not yet overwritten: step 5 only
                               // stores columns < c0 + 32) so that their global-memory latency hides behind the factorisation
      const int c1 = c0 + 32, m2 = d - c1, nc2 = min(32, m2);
      for (int e = tid - 32; e < m2 * 32; e += CB_THREADS - 32) {
        const int rr = e >> 5, cc = e & 31;
        const int r = c1 + rr, c = c1 + cc;
        double v = 0.0;
        if (c <= r && cc < nc2) {
          v = scale * src[rowoff(r) + c];
          if (c == r) v += pr[r];
        }
        pan2[rr * 33 + cc] = v;
      }
    }
    __syncthreads();
    // 4. rows below the block: x = a L_diag^-T, one thread per row (forward substitution against the 32 x 32 block)
    for (int rr = 32 + tid; rr < m; rr += CB_THREADS) {
      double x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = pan[rr * 33 + c];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        double s = x[c];
#pragma unroll
        for (int b = 0; b < c; ++b) s = fma(-x[b], pan[c * 33 + b], s);
        x[c] = s * red[c];
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) pan[rr * 33 + c] = x[c];
    }
    __syncthreads();
    // 5. write the factored panel back (packed lower)
    for (int e = tid; e < m * 32; e += CB_THREADS) {
      const int rr = e >> 5, cc = e & 31;
      const int r = c0 + rr, c = c0 + cc;
      if (c <= r && cc < nc) Lg[rowoff(r) + c] = pan[rr * 33 + cc];
    }
    __threadfence_block();
  }
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 1);
}


// ---- A2, second version: 16 warps, no barrier inside the panel update ------------------------------------------------
// Same left-looking blocked factorisation, reorganised around what the profile of the first version showed (one CTA of 8
// warps per SM, 86 % of the issue slots empty: every 32-column step of the update staged its B tile behind two barriers
// and waited for its A fragments from L2):
//   * the update  panel -= L[rows, <c0] L[c0:c0+32, <c0]'  runs WITHOUT barriers: every warp walks K = c0 in chunks of 8
//     columns on its own m-tiles, A fragments straight from L2 in fragment layout and fetched one chunk ahead into a second
//     register set, B fragments (the 32 rows of the diagonal block, all K = c0 columns: the same for all 16 warps) from a
//     shared-memory copy staged once per panel with a conflict-free pitch (B from L1 in fragment layout costs eight
//     L1 wavefronts per load and made the update L1-bound);
//   * the panel's source values are fetched by asynchronous copies issued BEFORE the update and only waited for when the
//     accumulators are folded into the panel, so their latency hides behind the update instead of a dedicated prefetch
//     buffer (half the shared memory);
//   * 16 warps per CTA: the row solves and the write-back use 512 threads.
// MTW = m-tiles (8 rows) per warp: ceil(ceil(d / 8) / 16).
constexpr int CB2_THREADS = 512, CB2_WARPS = 16, CB2_KC = 8;
// acc[i] += L[rows of m-tile (warp + 16 i), 0:c0] * B', i < NL: A fragments from global memory (row pointers Ar, lane offset
// fc folded in), fetched one 8-column chunk ahead; B fragments from the staged tile (Br: lane's row and column offset folded in).
template <int NL, int MTW>
__device__ __forceinline__ void chol_panel_update_impl(double (&acc)[MTW][4][2], const double* __restrict__ Lg, const double* __restrict__ Br,
                                                       const int pb, const int r0, const int d, const int c0, const int fc) {
  const double* Ar[NL];
  bool av[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const int r = r0 + 8 * CB2_WARPS * i;
    av[i] = r < d;
    Ar[i] = Lg + rowoff(min(r, d - 1)) + fc;
  }
  double a0[NL][2], a1[NL][2];
  auto load_a = [&](int kk, double (&a)[NL][2]) {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      a[i][0] = av[i] ? Ar[i][kk] : 0.0;
      a[i][1] = av[i] ? Ar[i][kk + 4] : 0.0;
    }
  };
  auto chunk = [&](int kk, const double (&a)[NL][2]) {
    double b[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      b[nt][0] = Br[8 * nt * pb + kk];
      b[nt][1] = Br[8 * nt * pb + kk + 4];
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int i = 0; i < NL; ++i)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[i][nt][0], acc[i][nt][1], a[i][ks], b[nt][ks]);
  };
  load_a(0, a0);
  for (int kk = 0; kk < c0; kk += 2 * CB2_KC) {  // c0 is a multiple of 32: an even number of chunks
    load_a(kk + CB2_KC, a1);
    chunk(kk, a0);
    if (kk + 2 * CB2_KC < c0) load_a(kk + 2 * CB2_KC, a0);
    chunk(kk + CB2_KC, a1);
  }
}
template <int NL, int MTW>
__device__ __forceinline__ void chol_panel_update(double (&acc)[MTW][4][2], const double* Lg, const double* Br, int pb, int r0, int d, int c0, int fc) {
  if constexpr (NL <= MTW) chol_panel_update_impl<NL, MTW>(acc, Lg, Br, pb, r0, d, c0, fc);
}
template <int MTW>
__global__ void __launch_bounds__(CB2_THREADS, 1) coef_chol_cta2_kernel(const Dims D, const DevPtrs P, double* Pfac) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const long long job = blockIdx.x;
  const int u = (int)(job / D.k), j = (int)(job % D.k), w = u / D.C;
  const int k = D.k, d = D.d;
  const bool SV = D.sv != 0;
  double* pan = sm;                   // [d][33] current panel, rows c0 .. d-1
  double* red = pan + (size_t)d * 33; // [64] scratch
  double* bsm = red + 64;             // [32][c0 + 4] rows c0 .. c0+31 of L, columns < c0
  __shared__ int bad;
  __shared__ double colbuf[64];
  if (tid == 0) bad = 0;
  double* Lg = Pfac + ((size_t)u * k + j) * D.ldS;  // in/out (SV: holds P_j on entry)
  const double* src = Lg;
  double scale = 1.0;
  if (!SV) {  // P_j = (sum_{i>=j} L_ij^2 / d_i) X'X
    double ws = 0.0;
    for (int i = j + tid; i < k; i += CB2_THREADS) {
      const double l = P.L[((size_t)u * k + i) * k + j];
      ws += l * l / P.diag[(size_t)u * k + i];
    }
    scale = block_sum(ws, red);
    src = P.XtX + (size_t)w * D.ldS;
  }
  const double* pr = P.prec + (size_t)u * k * d + (size_t)j * d;

  for (int c0 = 0; c0 < d; c0 += 32) {
    const int m = d - c0;            // rows of the panel
    const int nc = min(32, m);       // columns of the panel
    __syncthreads();                 // the previous panel is written back; pan is free
    // 1. source values of the panel (rows r = c0 + rr, columns c0 .. min(c0 + 31, r)): asynchronous copies, consumed in step 3
    for (int rr = warp; rr < m; rr += CB2_WARPS) {  // one warp per row: the packed row offset is computed once
      const int r = c0 + rr;
      if (lane < nc && c0 + lane <= r) cp_async8(pan + rr * 33 + lane, src + rowoff(r) + c0 + lane, 8);
      else pan[rr * 33 + lane] = 0.0;
    }
    cp_async_commit();
    const int pb = c0 + 4;  // pitch = 4 (mod 8): the B fragment loads of a half-warp hit 16 different banks
    for (int nn = warp; nn < 32; nn += CB2_WARPS) {
      const bool ok = c0 + nn < d;
      const double* row = Lg + rowoff(min(c0 + nn, d - 1));
      for (int kk = lane; kk < c0; kk += 32) {
        if (ok) cp_async8(bsm + nn * pb + kk, row + kk, 8);
        else bsm[nn * pb + kk] = 0.0;
      }
    }
    cp_async_commit();
    // 2. acc = L[rows, 0:c0] * L[c0:c0+32, 0:c0]'  (DMMA), no barriers inside
    double acc[MTW][4][2];
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
    const int mtiles = (m + 7) / 8;
    if (c0 > 0) {
      cp_async_wait<0>();
      __syncthreads();  // B tile (and the panel source) staged
      // m-tiles of this warp that exist in this panel (mt = warp + 16 i < mtiles): a prefix of its slots.  Dispatched to a
      // body with a compile-time count: dead slots issue nothing (the late panels have few rows and the longest K)
      const int nlive = mtiles > warp ? min(MTW, (mtiles - warp + CB2_WARPS - 1) / CB2_WARPS) : 0;
      const double* Br = bsm + fr * pb + fc;
      const int r0 = c0 + 8 * warp + fr;
      switch (nlive) {
        case 4: if (MTW >= 4) chol_panel_update<4>(acc, Lg, Br, pb, r0, d, c0, fc); break;
        case 3: if (MTW >= 3) chol_panel_update<3>(acc, Lg, Br, pb, r0, d, c0, fc); break;
        case 2: if (MTW >= 2) chol_panel_update<2>(acc, Lg, Br, pb, r0, d, c0, fc); break;
        case 1: chol_panel_update<1>(acc, Lg, Br, pb, r0, d, c0, fc); break;
        default: break;
      }
    } else {
      cp_async_wait<0>();
      __syncthreads();
    }
    // 3. panel = scale * source + prior diagonal - acc  (every row of the panel belongs to exactly one warp's m-tiles)
#pragma unroll
    for (int i = 0; i < MTW; ++i) {  // fragments: row 8 mt + fr, columns 8 nt + 2 fc + {0, 1}
      const int mt = warp + CB2_WARPS * i;
      if (mt >= mtiles) continue;
      const int rr = 8 * mt + fr;
      if (rr < m) {
        const double prr = pr[c0 + rr];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc = 8 * nt + 2 * fc + h;
            double v = scale * pan[rr * 33 + cc] - acc[i][nt][h];
            if (cc == rr) v += prr;
            pan[rr * 33 + cc] = v;
          }
        }
      }
    }
    __syncthreads();
    // 4. factor the diagonal block: one warp, the block in registers (one row per lane)
    if (warp == 0) {
      double dg[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) dg[c] = (lane < nc && c <= lane) ? pan[lane * 33 + c] : (c == lane ? 1.0 : 0.0);
      const double dinv = chol32_smem_chain(dg, colbuf, &bad);
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (lane < nc && c <= lane) pan[lane * 33 + c] = dg[c];
      red[lane] = dinv;  // 1 / L_cc for the row solves
    }
    __syncthreads();
    // 5. rows below the block: x = a L_diag^-T, one thread per row (forward substitution against the 32 x 32 block)
    for (int rr = 32 + tid; rr < m; rr += CB2_THREADS) {
      double x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = pan[rr * 33 + c];
#pragma unroll
      for (int c = 0; c < 32; ++c) {  // four partial sums: the dependent chain of a step is c / 4 FMAs, not c
        double s0 = x[c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int b = 0; b + 3 < c; b += 4) {
          s0 = fma(-x[b], pan[c * 33 + b], s0);
          s1 = fma(-x[b + 1], pan[c * 33 + b + 1], s1);
          s2 = fma(-x[b + 2], pan[c * 33 + b + 2], s2);
          s3 = fma(-x[b + 3], pan[c * 33 + b + 3], s3);
        }
#pragma unroll
        for (int b = c & ~3; b < c; ++b) s0 = fma(-x[b], pan[c * 33 + b], s0);
        x[c] = ((s0 + s1) + (s2 + s3)) * red[c];
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) pan[rr * 33 + c] = x[c];
    }
    __syncthreads();
    // 6. write the factored panel back (packed lower)
    for (int rr = warp; rr < m; rr += CB2_WARPS) {
      const int r = c0 + rr;
      if (lane < nc && c0 + lane <= r) Lg[rowoff(r) + c0 + lane] = pan[rr * 33 + lane];
    }
    __threadfence_block();
  }
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 1);
}

// ---- A3: invert the diagonal blocks in place, one warp per (matrix, block) ---------------------------------------
constexpr int ID_WARPS = 4;
__global__ void __launch_bounds__(32 * ID_WARPS) invert_diag_kernel(const Dims D, double* Pfac, long long nblocks_total, int nb) {
  __shared__ double blks[ID_WARPS][32 * 33];
  __shared__ double dinvs[ID_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long job = (long long)blockIdx.x * ID_WARPS + warp;
  if (job >= nblocks_total) return;
  const long long mat = job / nb;
  const int jb = (int)(job % nb), c0 = 32 * jb, nc = min(32, D.d - c0);
  double* Lg = Pfac + (size_t)mat * D.ldS;
  double* blk = blks[warp];
#pragma unroll 4
  for (int r = 0; r < 32; ++r)  // coalesced row segments of the packed factor
    blk[r * 33 + lane] = (r < nc && lane <= r) ? Lg[rowoff(c0 + r) + c0 + lane] : 0.0;
  __syncwarp();
  dinvs[warp][lane] = lane < nc ? 1.0 / blk[lane * 33 + lane] : 1.0;
  __syncwarp();
  double x[32];
  inv32_column([&](int r) { return blk + r * 33; }, dinvs[warp], lane, nc, x);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; ++r) blk[r * 33 + lane] = x[r];  // column `lane` of the inverse
  __syncwarp();
#pragma unroll 4
  for (int r = 0; r < 32; ++r)
    if (r < nc && lane <= r) Lg[rowoff(c0 + r) + c0 + lane] = blk[r * 33 + lane];
}

}  // namespace

static size_t chol_cta2_smem(const Dims& D) { return ((size_t)D.d * 33 + 64 + 32 * (size_t)(((D.d - 1) / 32) * 32 + 4)) * sizeof(double); }
// first version of the kernel: A/B switch (BVHAR_CHOL_V1), and designs whose B tile no longer fits beside the panel (d > 440)
static bool chol_cta_v1(const Dims& D) {
  static const bool v = getenv("BVHAR_CHOL_V1") != nullptr;
  return v || chol_cta2_smem(D) > 227 * 1024;
}
size_t chol_cta_smem(const Dims& D) {
  if (!chol_cta_v1(D)) return chol_cta2_smem(D);
  return ((size_t)D.d * 33 + 32 * CB_PB + 64 + (size_t)std::max(0, D.d - 32) * 33) * sizeof(double);
}
size_t seq_big_smem(const Dims& D) {
  return ((((size_t)D.k * (D.k + 1) / 2 + 1) & ~(size_t)1) + 3 * (size_t)D.d + 32 * 33 + (size_t)(CB_THREADS / 32) * D.d + D.k + 8) * sizeof(double);
}

bool coef_big_fits(const Dims& D, size_t smem_limit) {
  return D.d <= 512 && chol_cta_smem(D) <= smem_limit && (D.sv ? seq_time_smem(D) : seq_big_smem(D)) <= smem_limit;
}

cudaError_t launch_chol_big(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st) {
  const size_t smem = chol_cta_smem(D);
  if (smem > smem_limit) return cudaErrorInvalidConfiguration;
  const unsigned grid = (unsigned)((long long)D.U * D.k);
  if (!chol_cta_v1(D)) {
    const int mtw2 = ((D.d + 7) / 8 + CB2_WARPS - 1) / CB2_WARPS;
#define BV_CHOL2(MTWV)                                                                                         \
  do {                                                                                                         \
    cudaFuncSetAttribute(coef_chol_cta2_kernel<MTWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    coef_chol_cta2_kernel<MTWV><<<grid, CB2_THREADS, smem, st>>>(D, P, Pfac);                                  \
  } while (0)
    if (mtw2 <= 1) BV_CHOL2(1);
    else if (mtw2 <= 2) BV_CHOL2(2);
    else if (mtw2 <= 3) BV_CHOL2(3);
    else BV_CHOL2(4);
#undef BV_CHOL2
  } else {
  const int mtw = ((D.d + 7) / 8 + 7) / 8;
#define BV_CHOL(MTWV)                                                                                         \
  do {                                                                                                        \
    cudaFuncSetAttribute(coef_chol_cta_kernel<MTWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    coef_chol_cta_kernel<MTWV><<<grid, CB_THREADS, smem, st>>>(D, P, Pfac);                                   \
  } while (0)
  if (mtw <= 2) BV_CHOL(2);
  else if (mtw <= 4) BV_CHOL(4);
  else if (mtw <= 6) BV_CHOL(6);
  else BV_CHOL(8);
#undef BV_CHOL
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int nb = (D.d + 31) / 32;
  const long long nblk = (long long)D.U * D.k * nb;
  invert_diag_kernel<<<(unsigned)((nblk + ID_WARPS - 1) / ID_WARPS), 32 * ID_WARPS, 0, st>>>(D, Pfac, nblk, nb);
  return cudaGetLastError();
}

// part 1: batched Cholesky (one CTA per matrix), part 2: sequential draws.  (Part 0, the P build, is shared with
// the small-d path: launch_coef_v2(0, ...).)
cudaError_t launch_coef_big(int part, const Dims& D, const DevPtrs& P, double* Pfac, double* hscr, size_t smem_limit, cudaStream_t st,
                            const double* Rres) {
  if (part == 1) return launch_chol_big(D, P, Pfac, smem_limit, st);
  if (D.sv) return launch_seq_time(D, P, Pfac, hscr, smem_limit, st);  // SV: time-domain recurrence (kernels_seq_big.cu)
  const size_t smem = seq_big_smem(D);
  if (smem > smem_limit) return cudaErrorInvalidConfiguration;
  return (D.d + 31) / 32 <= 8 ? launch_seq_big_nc8(D, P, Pfac, hscr, smem, st, Rres) : launch_seq_big_nc16(D, P, Pfac, hscr, smem, st, Rres);
}

}  // namespace bvhar
