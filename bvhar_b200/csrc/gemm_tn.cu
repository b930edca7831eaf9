// gemm_tn.cu - the grouped fp64 "TN" contraction kernel (DMMA) declared in gemm_tn.cuh: G1 / G2 / G3 of the sweep, the
// X'X / X'Y set-up products of the LDLT models and the companion-matrix squarings of the stability filter.
#include <algorithm>
#include <cstdlib>

#include "gemm_tn.cuh"

#include <cstdlib>

namespace bvhar {

// Loads a GBK x 128 tile (rows k0.., cols c0..) of a row-major matrix into smem [GBK][128 + GPAD].
template <bool ALIGNED16>
__device__ __forceinline__ void gemm_load_tile(double* sm, const double* __restrict__ g, int ld, int k0, int c0, int K,
                                               int ncols, int tid) {
  if (ALIGNED16) {
#pragma unroll
    for (int i = 0; i < (GBK * 64) / GEMM_THREADS; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int row = c >> 6, col = (c & 63) << 1;
      const bool ok = (k0 + row < K) && (c0 + col < ncols);  // ncols rounded up to even by the caller
      const double* src = ok ? g + (size_t)(k0 + row) * ld + c0 + col : g;
      cp_async16(sm + row * (128 + GPAD) + col, src, ok ? 16 : 0);
    }
  } else {
#pragma unroll
    for (int i = 0; i < (GBK * 128) / GEMM_THREADS; ++i) {
      const int c = tid + i * GEMM_THREADS;
      const int row = c >> 7, col = c & 127;
      const bool ok = (k0 + row < K) && (c0 + col < ncols);
      const double* src = ok ? g + (size_t)(k0 + row) * ld + c0 + col : g;
      cp_async8(sm + row * (128 + GPAD) + col, src, ok ? 8 : 0);
    }
  }
}

// One k-tile of the warp's 64 x 32 block: m-tiles 2 i + wr (wr = warp row 0 | 1, interleaved so that a partial row
// tile - e.g. rows 128..199 of the M = C k = 200 rows of a 4-chain window at k = 50 - splits its valid m-tiles evenly
// over the two warp rows), n-tiles j.  NMT = m-tiles this warp owns, a compile-time count: a row tile with 72 valid rows
// issues 5/8 of the DMMAs of a full one (predicated-off DMMAs still occupy the tensor pipe, so the count must be static).
template <int NMT>
__device__ __forceinline__ void gemm_ktile(double (&acc)[8][4][2], const double* as, const double* bs, int wr, int wn0, int fr, int fc) {
#pragma unroll
  for (int kk = 0; kk < GBK; kk += 4) {
    double af[NMT], bf[4];
#pragma unroll
    for (int i = 0; i < NMT; ++i) af[i] = as[(kk + fc) * (GBM + GPAD) + (2 * i + wr) * 8 + fr];
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * (GBN + GPAD) + wn0 + j * 8 + fr];
#pragma unroll
    for (int i = 0; i < NMT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

// PART = false: every row tile is full (M a multiple of 128, e.g. the 1 024-chain C3 batch): the plain loop, m-tiles
// wr * 8 + i.  PART = true: row tiles may be partial: interleaved m-tiles and a static per-warp m-tile count.
template <bool ALIGNED16, bool PART>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_tn_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  double* As = reinterpret_cast<double*>(gemm_smem_raw);
  double* Bs = As + GSTAGES * GBK * (GBM + GPAD);
  const int grp = blockIdx.z;
  const int K = g.Kg ? g.Kg[grp] : g.K;
  const int M = g.Mg ? g.Mg[grp] : g.M;
  if ((int)blockIdx.y * GBM >= M) return;
  const double* A = g.A + (g.a_off ? g.a_off[grp] : 0);
  const double* B = g.B + (g.b_off ? g.b_off[grp] : 0);
  double* C = g.C + (g.c_off ? g.c_off[grp] : 0);
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 2, wn0 = (warp & 3) * 32;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row (0..7) / k index (0..3)
  const int Mld = ALIGNED16 ? ((M + 1) & ~1) : M;
  const int Nld = ALIGNED16 ? ((g.N + 1) & ~1) : g.N;
  const int mt_valid = m0 + GBM > M ? (M - m0 + 7) / 8 : 16;  // valid m-tiles of this row tile
  const int nmt = PART ? (mt_valid - wr + 1) / 2 : 8;         // ... of which this warp row owns 2 i + wr < mt_valid

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + GBK - 1) / GBK;
#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) {
      gemm_load_tile<ALIGNED16>(As + s * GBK * (GBM + GPAD), A, g.lda, s * GBK, m0, K, Mld, tid);
      gemm_load_tile<ALIGNED16>(Bs + s * GBK * (GBN + GPAD), B, g.ldb, s * GBK, n0, K, Nld, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot freed at iteration kt - 1
      const int nx = kt + GSTAGES - 1;
      if (nx < nk) {
        const int s = nx % GSTAGES;
        gemm_load_tile<ALIGNED16>(As + s * GBK * (GBM + GPAD), A, g.lda, nx * GBK, m0, K, Mld, tid);
        gemm_load_tile<ALIGNED16>(Bs + s * GBK * (GBN + GPAD), B, g.ldb, nx * GBK, n0, K, Nld, tid);
      }
      cp_async_commit();
    }
    const double* as = As + (kt % GSTAGES) * GBK * (GBM + GPAD);
    const double* bs = Bs + (kt % GSTAGES) * GBK * (GBN + GPAD);
    if (!PART) {
#pragma unroll
      for (int kk = 0; kk < GBK; kk += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) af[i] = as[(kk + fc) * (GBM + GPAD) + wr * 64 + i * 8 + fr];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * (GBN + GPAD) + wn0 + j * 8 + fr];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
      switch (nmt) {  // warp-uniform
        case 8: gemm_ktile<8>(acc, as, bs, wr, wn0, fr, fc); break;
        case 7: gemm_ktile<7>(acc, as, bs, wr, wn0, fr, fc); break;
        case 6: gemm_ktile<6>(acc, as, bs, wr, wn0, fr, fc); break;
        case 5: gemm_ktile<5>(acc, as, bs, wr, wn0, fr, fc); break;
        case 4: gemm_ktile<4>(acc, as, bs, wr, wn0, fr, fc); break;
        case 3: gemm_ktile<3>(acc, as, bs, wr, wn0, fr, fc); break;
        case 2: gemm_ktile<2>(acc, as, bs, wr, wn0, fr, fc); break;
        case 1: gemm_ktile<1>(acc, as, bs, wr, wn0, fr, fc); break;
        default: break;
      }
    }
  }
  cp_async_wait<0>();

  const long long yoff = (g.epi == EPI_Y_MINUS && g.y_off) ? g.y_off[grp] : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (PART ? (2 * i + wr) * 8 : wr * 64 + i * 8) + fr;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wn0 + j * 8 + fc * 2;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (g.epi == EPI_Y_MINUS) {
        if (col < g.N) v0 = g.Y[yoff + (size_t)row * g.ldy + (col % g.ymod)] - v0;
        if (col + 1 < g.N) v1 = g.Y[yoff + (size_t)row * g.ldy + ((col + 1) % g.ymod)] - v1;
      }
      double* dst = C + (size_t)row * g.ldc + col;
      if (col + 1 < g.N && ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      } else {
        if (col < g.N) dst[0] = v0;
        if (col + 1 < g.N) dst[1] = v1;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// The Blackwell form of the same contraction (the default whenever the operands are 16-byte addressable):
//   * PERSISTENT: one CTA per SM walks the tile list (group, row tile, column tile), so the operand pipeline never
//     drains between tiles - the producer is already filling the first stages of the next tile while the consumers
//     write the previous tile's 128 x 128 block;
//   * WARP-SPECIALISED: warp 8 is the producer and issues nothing but bulk asynchronous copies (the TMA unit,
//     cp.async.bulk.shared.global with mbarrier complete_tx: UBLKCP + SYNCS in the SASS) - one 1 KB row of the A or B
//     tile per lane, into the same padded rows the fragment loads want (conflict-free, no swizzle needed); warps 0-7
//     only load fragments and issue DMMAs.  No thread of the consumer warps computes a global address in the loop;
//   * mbarrier pipeline, 4 stages of 33 KB: full[s] (producer arrival + transaction bytes) / empty[s] (one arrival
//     per consumer warp).  There is no __syncthreads in the loop: the eight consumer warps drift apart, and while one
//     waits on a barrier the other warp of its scheduler keeps the DMMA pipe fed.
// Rows past K of the last k-tile are zero-filled by the producer lanes (generic stores, ordered before the arrival
// by __syncwarp + the release of the arrive, and before later async writes by fence.proxy.async).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TB_STAGES = 4;
constexpr int TB_CONSUMERS = 8;
constexpr int TB_THREADS = (TB_CONSUMERS + 1) * 32;
constexpr int TB_ROW = GBM + GPAD;                        // padded smem row (doubles); GBM == GBN
constexpr int TB_STAGE_DOUBLES = 2 * GBK * TB_ROW;        // A rows then B rows
constexpr size_t TB_SMEM = (size_t)TB_STAGES * TB_STAGE_DOUBLES * sizeof(double) + 2 * TB_STAGES * sizeof(uint64_t);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TileCoord { int grp, m0, n0, M, K, sub, k0, slice; };
// Work item w of the persistent schedule -> (group, row tile, column tile), column tiles fastest: the CTAs running at
// one time share a few A row panels and the whole (L2-resident) B.  Items below `split_from` are whole 128 x 128
// tiles.  The tiles of the last, partial round (2 400 tiles on 148 SMs = 16 full rounds + 32 tiles) would leave most
// SMs idle for a whole tile time, so they are cut into four 128 x 32 column quarters (sub = 0..3) that spread over
// all the CTAs.  Returns 0 = past the end, 1 = skip (row tile past the group's M / quarter past N), 2 = work.
__device__ __forceinline__ int work_decode(const GemmArgs& g, int w, int mtiles, int ntiles, int total, int split_from, TileCoord& t) {
  int tile = w;
  t.sub = -1;
  if (w >= split_from) {
    const int q = w - split_from;
    tile = split_from + (q >> 2);
    t.sub = q & 3;
  }
  if (tile >= total) return 0;
  t.slice = 0;
  if (g.ksplit > 1) {  // K slices fastest: the slices of one output tile run side by side on different SMs
    t.slice = tile % g.ksplit;
    tile /= g.ksplit;
  }
  const int nt = tile % ntiles, rest = tile / ntiles;
  const int mt = rest % mtiles;
  t.grp = rest / mtiles;
  t.K = g.Kg ? g.Kg[t.grp] : g.K;
  t.k0 = 0;
  if (g.ksplit > 1) {  // slice = a whole number of k-tiles; trailing slices may be empty (they store zeros)
    const int per = ((t.K + GBK - 1) / GBK + g.ksplit - 1) / g.ksplit * GBK;
    t.k0 = t.slice * per;
    t.K = max(0, min(per, t.K - t.k0));
  }
  t.M = g.Mg ? g.Mg[t.grp] : g.M;
  t.m0 = mt * GBM;
  t.n0 = nt * GBN + (t.sub > 0 ? t.sub * 32 : 0);
  return (t.m0 < t.M && t.n0 < g.N) ? 2 : 1;
}

// NARROW (N <= 64, full row tiles): the eight consumer warps form a 4 x 2 grid (32 rows x 32 columns each) instead of
// 2 x 4: with the wide layout a right-hand side of 61 columns leaves the warp columns 2 and 3 - and with them two of the
// SM's four tensor sub-partitions - without a single useful DMMA.
template <bool PART, bool NARROW = false>
__global__ void __launch_bounds__(TB_THREADS, 1) dgemm_tn_bulk_kernel(const GemmArgs g, const int mtiles, const int ntiles, const int total,
                                                                      const int split_from) {
  extern __shared__ __align__(128) unsigned char gemm_smem_raw[];
  double* stages = reinterpret_cast<double*>(gemm_smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(gemm_smem_raw + (size_t)TB_STAGES * TB_STAGE_DOUBLES * sizeof(double));
  uint64_t* empty = full + TB_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TB_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, TB_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  const int Nld = (g.N + 1) & ~1;
  unsigned it = 0;  // k-tiles handled so far by this CTA: stage = it % STAGES, phase = it / STAGES (both roles count alike)

  if (warp == TB_CONSUMERS) {
    // ---------------- producer warp ----------------
    const int r = lane & 15;
    const bool isB = lane >= 16;
    for (int w = blockIdx.x;; w += gridDim.x) {
      TileCoord t;
      const int kind = work_decode(g, w, mtiles, ntiles, total, split_from, t);
      if (kind == 0) break;
      if (kind == 1) continue;
      const int Mld = (t.M + 1) & ~1;
      const unsigned bytesA = (unsigned)min(GBM, Mld - t.m0) * 8u, bytesB = (unsigned)min(t.sub >= 0 ? 32 : GBN, Nld - t.n0) * 8u;
      const double* src = isB ? g.B + (g.b_off ? g.b_off[t.grp] : 0) + t.n0 : g.A + (g.a_off ? g.a_off[t.grp] : 0) + t.m0;
      const size_t ld = isB ? g.ldb : g.lda;
      const unsigned mybytes = isB ? bytesB : bytesA;
      const int nk = (t.K + GBK - 1) / GBK;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % TB_STAGES;
        if (it >= TB_STAGES) mbar_wait(empty + s, ((it / TB_STAGES) & 1) ^ 1);
        const int k0 = kt * GBK, rows = min(GBK, t.K - k0);
        double* dst = stages + (size_t)s * TB_STAGE_DOUBLES + (isB ? GBK * TB_ROW : 0) + r * TB_ROW;
        if (rows < GBK) {  // ragged last k-tile: the rows past K must contribute nothing
          if (r >= rows)
            for (int c = 0; c < TB_ROW; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(0.0, 0.0);
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(full + s, (unsigned)rows * (bytesA + bytesB));
        __syncwarp();
        if (r < rows) bulk_copy_g2s(dst, src + (size_t)(t.k0 + k0 + r) * ld, mybytes, full + s);
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  constexpr int NMT = NARROW ? 4 : 8, WROWS = NARROW ? 32 : 64;  // m-tiles and rows per warp (full-tile path)
  const int wr = NARROW ? warp >> 1 : warp >> 2, wn0 = NARROW ? (warp & 1) * 32 : (warp & 3) * 32;
  const int fr = lane >> 2, fc = lane & 3;
  for (int w = blockIdx.x;; w += gridDim.x) {
    TileCoord t;
    const int kind = work_decode(g, w, mtiles, ntiles, total, split_from, t);
    if (kind == 0) break;
    if (kind == 1) continue;
    const int M = t.M, m0 = t.m0, n0 = t.n0;
    double* C = (g.ksplit > 1 ? g.ws + (size_t)t.slice * g.ws_stride : g.C) + (g.c_off ? g.c_off[t.grp] : 0);
    const long long yoff = (g.epi == EPI_Y_MINUS && g.y_off) ? g.y_off[t.grp] : 0;
    // (ym0, ym1: col % ymod and (col + 1) % ymod, computed once per column pair of the tile - the runtime modulo cost as
    // much as the rest of the epilogue when it was evaluated per stored element)
    auto store2 = [&](int row, int col, double v0, double v1, int ym0, int ym1) {
      if (g.epi == EPI_Y_MINUS) {
        if (col < g.N) v0 = g.Y[yoff + (size_t)row * g.ldy + ym0] - v0;
        if (col + 1 < g.N) v1 = g.Y[yoff + (size_t)row * g.ldy + ym1] - v1;
      }
      double* dst = C + (size_t)row * g.ldc + col;
      if (col + 1 < g.N && ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      } else {
        if (col < g.N) dst[0] = v0;
        if (col + 1 < g.N) dst[1] = v1;
      }
    };
    const int nk = (t.K + GBK - 1) / GBK;

    if (t.sub >= 0) {
      // a 128 x 32 quarter of a last-round tile: warp w owns rows 16 w .. 16 w + 15 (2 m-tiles) x 4 n-tiles
      double qa[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) qa[i][j][0] = qa[i][j][1] = 0.0;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % TB_STAGES;
        mbar_wait(full + s, (it / TB_STAGES) & 1);
        const double* as = stages + (size_t)s * TB_STAGE_DOUBLES;
        const double* bs = as + GBK * TB_ROW;
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
          double af[2], bf[4];
#pragma unroll
          for (int i = 0; i < 2; ++i) af[i] = as[(kk + fc) * TB_ROW + warp * 16 + i * 8 + fr];
#pragma unroll
          for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * TB_ROW + j * 8 + fr];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(qa[i][j][0], qa[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
      }
      const int qbase = g.epi == EPI_Y_MINUS ? (n0 + fc * 2) % g.ymod : 0;  // one runtime modulo per tile; the columns advance by 8
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = m0 + warp * 16 + i * 8 + fr;
        if (row >= M) continue;
        int cm = qbase;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          store2(row, n0 + j * 8 + fc * 2, qa[i][j][0], qa[i][j][1], cm, cm + 1 == g.ymod ? 0 : cm + 1);
          cm += 8;
          while (cm >= g.ymod && g.epi == EPI_Y_MINUS) cm -= g.ymod;
        }
      }
      continue;
    }

    const int mt_valid = m0 + GBM > M ? (M - m0 + 7) / 8 : 16;
    const int nmt = PART ? (mt_valid - wr + 1) / 2 : 8;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // a warp whose 32 columns lie past N (narrow right-hand sides: N = 61 leaves two of the four warp columns empty) only
    // keeps the barrier protocol going: its DMMAs would take the tensor pipe from the warps that have columns.  (Its own
    // loop: a branch inside the hot loop below cost the full-width contractions 3 - 8 %.)
    if (n0 + wn0 >= g.N) {
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % TB_STAGES;
        mbar_wait(full + s, (it / TB_STAGES) & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
      }
      continue;
    }
    for (int kt = 0; kt < nk; ++kt, ++it) {
      const int s = it % TB_STAGES;
      mbar_wait(full + s, (it / TB_STAGES) & 1);
      const double* as = stages + (size_t)s * TB_STAGE_DOUBLES;
      const double* bs = as + GBK * TB_ROW;
      if (!PART) {
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
          double af[NMT], bf[4];
#pragma unroll
          for (int i = 0; i < NMT; ++i) af[i] = as[(kk + fc) * TB_ROW + wr * WROWS + i * 8 + fr];
#pragma unroll
          for (int j = 0; j < 4; ++j) bf[j] = bs[(kk + fc) * TB_ROW + wn0 + j * 8 + fr];
#pragma unroll
          for (int i = 0; i < NMT; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else {
        switch (nmt) {  // warp-uniform
          case 8: gemm_ktile<8>(acc, as, bs, wr, wn0, fr, fc); break;
          case 7: gemm_ktile<7>(acc, as, bs, wr, wn0, fr, fc); break;
          case 6: gemm_ktile<6>(acc, as, bs, wr, wn0, fr, fc); break;
          case 5: gemm_ktile<5>(acc, as, bs, wr, wn0, fr, fc); break;
          case 4: gemm_ktile<4>(acc, as, bs, wr, wn0, fr, fc); break;
          case 3: gemm_ktile<3>(acc, as, bs, wr, wn0, fr, fc); break;
          case 2: gemm_ktile<2>(acc, as, bs, wr, wn0, fr, fc); break;
          case 1: gemm_ktile<1>(acc, as, bs, wr, wn0, fr, fc); break;
          default: break;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    const int ybase = g.epi == EPI_Y_MINUS ? (n0 + wn0 + fc * 2) % g.ymod : 0;  // one runtime modulo per tile; the columns advance by 8
#pragma unroll
    for (int i = 0; i < (PART ? 8 : NMT); ++i) {
      const int row = m0 + (PART ? (2 * i + wr) * 8 : wr * WROWS + i * 8) + fr;
      if (row >= M) continue;
      int cm = ybase;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        store2(row, n0 + wn0 + j * 8 + fc * 2, acc[i][j][0], acc[i][j][1], cm, cm + 1 == g.ymod ? 0 : cm + 1);
        cm += 8;
        while (cm >= g.ymod && g.epi == EPI_Y_MINUS) cm -= g.ymod;
      }
    }
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double* __restrict__ ws, const int S, const long long stride, double* __restrict__ C,
                                                            const long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < S; ++q) s += ws[(size_t)q * stride + i];  // fixed order: bit-reproducible
    C[i] = s;
  }
}

// BVHAR_GEMM_LEGACY=1 keeps the sm_80-style cp.async kernel for every shape (A/B comparisons in the profiles).
static bool gemm_force_legacy() {
  static const bool v = [] { const char* e = getenv("BVHAR_GEMM_LEGACY"); return e && e[0] == '1'; }();
  return v;
}

int gemm_choose_ksplit(long long tiles, int kmax, int sms) {
  static const bool off = [] {
    const char* e = getenv("BVHAR_GEMM_NO_SPLITK");
    const char* l = getenv("BVHAR_GEMM_LEGACY");  // the cp.async kernel has no slices
    return (e && e[0] == '1') || (l && l[0] == '1');
  }();
  if (off || tiles <= 0 || 2 * tiles > sms) return 1;  // at least half a wave of tiles: no split
  const int by_sms = (int)(sms / tiles), by_k = kmax / (8 * GBK);  // keep >= 8 k-tiles per slice
  const int s = std::min(std::min(by_sms, by_k), 16);
  return s < 2 ? 1 : s;
}
cudaError_t launch_splitk_reduce(const GemmArgs& g, long long count, cudaStream_t st) {
  if (g.ksplit <= 1) return cudaSuccess;
  const int blocks = (int)std::min<long long>((count + 255) / 256, 148 * 8);
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g.ws, g.ksplit, g.ws_stride, g.C, count);
  return cudaGetLastError();
}

cudaError_t launch_dgemm_tn(const GemmArgs& g, cudaStream_t st) {
  const bool part = g.Mg != nullptr || (g.M % GBM) != 0;
  if (g.aligned16 && !gemm_force_legacy()) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(dgemm_tn_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
      cudaFuncSetAttribute(dgemm_tn_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
      cudaFuncSetAttribute(dgemm_tn_bulk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
    }
    const int mtiles = (g.M + GBM - 1) / GBM, ntiles = (g.N + GBN - 1) / GBN;
    const long long total = (long long)mtiles * ntiles * g.groups * (g.ksplit > 1 ? g.ksplit : 1);
    const int grid = (int)(total < sms ? total : sms);
    // tiles of the last partial round are split into column quarters when that round would fill at most half the SMs
    const long long rem = total % grid;
    const long long split_from = (total > grid && rem > 0 && 2 * rem <= grid) ? total - rem : total;
    if (part) dgemm_tn_bulk_kernel<true><<<grid, TB_THREADS, TB_SMEM, st>>>(g, mtiles, ntiles, (int)total, (int)split_from);
    else if (g.N <= 64) dgemm_tn_bulk_kernel<false, true><<<grid, TB_THREADS, TB_SMEM, st>>>(g, mtiles, ntiles, (int)total, (int)split_from);
    else dgemm_tn_bulk_kernel<false><<<grid, TB_THREADS, TB_SMEM, st>>>(g, mtiles, ntiles, (int)total, (int)split_from);
    return cudaGetLastError();
  }
  dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, g.groups);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dgemm_tn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    cudaFuncSetAttribute(dgemm_tn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    configured = true;
  }
  if (g.aligned16) {
    if (part) dgemm_tn_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
    else dgemm_tn_kernel<true, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
  } else {
    if (part) dgemm_tn_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
    else dgemm_tn_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
