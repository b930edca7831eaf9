// kernels_impact.cu — stage 7 (+8): updateImpact (triangular.h:322-336 + draw_coef draw.h:372-383 + draw_savs
// draw.h:398-415) and updateChol (triangular.h:348, build_inv_lower draw.h:124-132).
//
// Row j of L regresses Z_j / s_j on Z_{<j} / s_j with prior N(0, diag(1 / prec)).  The rows are conditionally
// independent given Z, so the work splits into
//
//   phase 1  all the time contractions of the unit at once.  With p = (a, b), b <= a, indexing the
//            k(k+1)/2 column pairs of Z and w_tj = 1 / s_tj^2:
//                T[p][j] = sum_t (z_ta z_tb) w_tj
//            holds the Gram of row j (entries with a < j) and its right-hand side (a == j).  In SV models this
//            is a [pairs x n] x [n x k] contraction per unit: the fp64 tensor pipe does it (DMMA
//            mma.sync.m8n8k4.f64, tcgen05 has no f64 kind) with the A fragment z_ta * z_tb formed on the fly
//            from a shared-memory tile of Z, so the pair products never touch memory.  Only tiles with
//            j >= a are issued.  LDLT models have w constant in t: one plain Z'Z per unit, computed by the grouped
//            DMMA contraction (engine: gram_Z) as a full k x kz matrix; updateState (draw.h:1244-1259) then needs
//            no pass over time either: ||(Z L')_i||^2 = L_i (Z'Z) L_i'.
//   phase 2  one warp per row: packed Cholesky of the j x j precision, the two triangular solves, the draw,
//            the SAVS copy, and the rebuilt row of L.
#include <algorithm>
#include <cstdlib>

#include "block_linalg.cuh"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"
#include "rng.cuh"

namespace bvhar {

namespace {

__device__ __forceinline__ int impact_row_off(int j) {  // sum_{j'<j} (j'(j'+1)/2 + j')
  return ((j - 1) * j * (2 * j - 1) / 6 + 3 * (j - 1) * j / 2) / 2;
}
__device__ __forceinline__ void pair_of(int pid, int& a, int& b) {  // pid = a(a+1)/2 + b, b <= a
  a = (int)((sqrt(8.0 * pid + 1.0) - 1.0) * 0.5);
  while (tri_idx(a + 1, 0) <= pid) ++a;
  while (tri_idx(a, 0) > pid) --a;
  b = pid - tri_idx(a, 0);
}

constexpr int IMP_TT_SMALL = 64, IMP_TT_LARGE = 32;  // time points per staged tile: 64 when the tile is narrow (k <= 32), else 32
__host__ __device__ constexpr int imp_tt(int nt) { return nt <= 4 ? IMP_TT_SMALL : IMP_TT_LARGE; }
constexpr int IMP_STAGES = 3;   // cp.async pipeline depth
constexpr int IMP_WARPS = 8;

// ---- phase 2, shared by both covariance models ---------------------------------------------------------
template <bool SV>
__device__ void impact_draw_rows(const Dims& D, const DevPtrs& P, const double* G, double* wbuf, int u, int* bad, int nw, int kz) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // nw = warps that own a workspace in wbuf
  const int k = D.k, q = D.q;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const int wstride = k * (k + 1) / 2 + 3 * k;
  double* Pw = wbuf + (size_t)warp * wstride;
  double* r = Pw + k * (k + 1) / 2;
  double* z = r + k;
  double* x = z + k;
  double* Lu = P.L + (size_t)u * k * k;
  for (int b = tid; b < k; b += blockDim.x) Lu[b] = b == 0 ? 1.0 : 0.0;  // row 0 of L (build_inv_lower, draw.h:124-132)
  if (warp >= nw) return;
  for (int j = 1 + warp; j < k; j += nw) {
    const int id = j * (j - 1) / 2;
    const double inv_dj = SV ? 1.0 : 1.0 / P.diag[(size_t)u * k + j];
    const double* Gj = SV ? G + impact_row_off(j) : G;
    if (SV) {
      for (int p = lane; p < j * (j + 1) / 2; p += 32) Pw[p] = Gj[p];
    } else {  // leading j x j block of the full Z'Z (row pitch kz)
      for (int a = 0; a < j; ++a)
        for (int b = lane; b <= a; b += 32) Pw[tri_idx(a, 0) + b] = G[a * kz + b] * inv_dj;
    }
    __syncwarp();
    for (int b = lane; b < j; b += 32) {
      const double pr = P.chol_prec[(size_t)u * q + id + b];
      Pw[tri_idx(b, b)] += pr;
      r[b] = SV ? Gj[j * (j + 1) / 2 + b] : G[j * kz + b] * inv_dj;  // prior_chol_mean = 0 (tri.h:52)
      z[b] = rng.normal_elem(ST_IMPACT_Z, (uint32_t)(id + b));
    }
    __syncwarp();
    chol_packed_warp(Pw, j, bad);
    precision_solve_warp(Pw, j, r, z, x);
    __syncwarp();
    for (int b = lane; b < k; b += 32) {
      double val = b == j ? 1.0 : 0.0;
      if (b < j) {
        val = x[b];
        P.contem[(size_t)u * q + id + b] = val;
        const double fit = fabs(val) * P.znorm[(size_t)u * k + b];  // draw_savs, draw.h:398-415
        P.sparse_contem[(size_t)u * q + id + b] = val * fmax(fit - 1.0 / (val * val), 0.0) / fit;
      }
      Lu[j * k + b] = val;
    }
    __syncwarp();
  }
}

// ---- SV: DMMA contraction + row draws, one CTA per unit ---------------------------------------------------
// NT = number of 8-wide tiles over the equation index j (k <= 8 NT); IMP_MT = m-tiles (8 pairs each) per warp
// and pass (accumulators: IMP_MT * NT * 2 doubles per thread).
//
// The staged tiles have the compile-time row pitch KPT = 8 NT + 4 doubles (= 4 mod 8: the fragment loads of a
// half-warp - four consecutive columns on four consecutive rows - hit 16 different banks; rows stay 16-byte aligned
// for the 16-byte asynchronous copies), so every fragment address is a per-tile register plus an immediate and a
// k-step is 2 LDS + 1 DMUL per m-tile and 1 LDS + 1 DMMA per issued (m-tile, j-tile).  An m-tile whose smallest a is in
// j-tile c only needs the j-tiles >= c (j >= a): the count of issued j-tiles is a warp-uniform property of the slot,
// dispatched ONCE per staged tile to a fully unrolled body with a compile-time count (the per-DMMA predicates of the
// first version cost 15 issued instructions per DMMA).
template <int NT, int C>
__device__ __forceinline__ void impact_slot_steps(double (*acc)[2], const double* __restrict__ za, const double* __restrict__ zb,
                                                  const double* __restrict__ dr) {
  constexpr int KPT = 8 * NT + 4, IMP_TT = imp_tt(NT);
#pragma unroll
  for (int s = 0; s < IMP_TT / 4; ++s) {
    const double af = za[s * 4 * KPT] * zb[s * 4 * KPT];
#pragma unroll
    for (int c = 0; c < C; ++c) dmma_m8n8k4(acc[NT - C + c][0], acc[NT - C + c][1], af, dr[s * 4 * KPT + 8 * (NT - C + c)]);
  }
}
template <int NT, int C>
struct ImpactSlot {
  static __device__ __forceinline__ void run(int n, double (*acc)[2], const double* za, const double* zb, const double* dr) {
    if (n == C) impact_slot_steps<NT, C>(acc, za, zb, dr);
    else ImpactSlot<NT, C - 1>::run(n, acc, za, zb, dr);
  }
};
template <int NT>
struct ImpactSlot<NT, 0> {
  static __device__ __forceinline__ void run(int, double (*)[2], const double*, const double*, const double*) {}
};

template <int NT, int IMP_MT>
__global__ void __launch_bounds__(IMP_WARPS * 32) impact_sv_kernel(const Dims D, const DevPtrs P, double* gscratch, const int gsize, const int pw) {
  extern __shared__ __align__(16) double sm[];
  constexpr int KPT = 8 * NT + 4, IMP_TT = imp_tt(NT);
  constexpr int NCOL = (8 * NT + 31) / 32;  // columns of Z per lane in the norm accumulation
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int k = D.k, ldM = D.ldM;
  const int n = P.win_n[w];
  const long long off = P.win_off[w] + (long long)c * k;
  double* zt = sm;                                   // [STAGES][TT][KPT]
  double* dt = zt + IMP_STAGES * IMP_TT * KPT;       // [STAGES][TT][KPT]
  double* nrm = dt + IMP_STAGES * IMP_TT * KPT;      // [k] squared column norms of Z
  double* wbuf = nrm + ((k + 1) & ~1);               // phase-2 per-warp workspaces
  double* G = gscratch + (size_t)u * gsize;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  const int npairs = k * (k + 1) / 2;
  const int mtiles = (npairs + 7) / 8;
  const int ntiles = (n + IMP_TT - 1) / IMP_TT;
  const int per_pass = IMP_MT * IMP_WARPS;
  double nacc[NCOL];
#pragma unroll
  for (int q = 0; q < NCOL; ++q) nacc[q] = 0.0;
  // the columns k .. KPT-1 of the staged tiles are never written by the copies: clear them once (they only reach
  // accumulator columns j >= k, which are not stored, but must not hold signalling garbage)
  for (int e = tid; e < 2 * IMP_STAGES * IMP_TT * KPT; e += IMP_WARPS * 32) sm[e] = 0.0;

  // staging: thread = (row within a group of rpi rows, chunk of the row); 16-byte chunks when the unit's columns are 16-byte aligned
  const bool al16 = ((k & 1) == 0) && ((off & 1) == 0) && ((ldM & 1) == 0);
  const int cw = al16 ? 2 : 1, kc = k / cw, rpi = (IMP_WARPS * 32) / kc;
  const int st_r = tid / kc, st_m = (tid - st_r * kc) * cw;
  auto stage_load = [&](int tile, int slot) {
    const int t0 = tile * IMP_TT;
    double* zs = zt + slot * IMP_TT * KPT;
    double* ds = dt + slot * IMP_TT * KPT;
    if (st_r < rpi) {
      for (int tt = st_r; tt < IMP_TT; tt += rpi) {
        const bool ok = t0 + tt < n;
        const long long at = ok ? off + (long long)(t0 + tt) * ldM + st_m : off;
        if (al16) {
          cp_async16(zs + tt * KPT + st_m, P.Z + at, ok ? 16 : 0);
          cp_async16(ds + tt * KPT + st_m, P.Dinv + at, ok ? 16 : 0);
        } else {
          cp_async8(zs + tt * KPT + st_m, P.Z + at, ok ? 8 : 0);
          cp_async8(ds + tt * KPT + st_m, P.Dinv + at, ok ? 8 : 0);
        }
      }
    }
  };

  for (int mt0 = 0; mt0 < mtiles; mt0 += per_pass) {
    // this warp's m-tiles of the pass: mt0 + warp + 8 i (heavy early tiles and light late tiles interleave)
    int pa[IMP_MT], pb[IMP_MT], nact[IMP_MT];
#pragma unroll
    for (int i = 0; i < IMP_MT; ++i) {
      const int mt = mt0 + warp + IMP_WARPS * i;
      int a0, b0;
      pair_of(min(8 * mt, npairs - 1), a0, b0);
      nact[i] = mt < mtiles ? NT - (a0 >> 3) : 0;  // j-tiles below the smallest a of the m-tile are never needed (j >= a)
      pair_of(min(8 * mt + fr, npairs - 1), pa[i], pb[i]);
    }
    double acc[IMP_MT][NT][2];
#pragma unroll
    for (int i = 0; i < IMP_MT; ++i)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;

    __syncthreads();  // the previous pass is done with the staging buffers (first pass: the clear above)
#pragma unroll
    for (int s = 0; s < IMP_STAGES - 1; ++s) {
      if (s < ntiles) stage_load(s, s);
      cp_async_commit();
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      cp_async_wait<IMP_STAGES - 2>();
      __syncthreads();
      {
        const int nx = tile + IMP_STAGES - 1;
        if (nx < ntiles) stage_load(nx, nx % IMP_STAGES);
        cp_async_commit();
      }
      const double* zs = zt + (tile % IMP_STAGES) * IMP_TT * KPT + fc * KPT;
      const double* ds = dt + (tile % IMP_STAGES) * IMP_TT * KPT + fc * KPT + fr;
      if (mt0 == 0) {  // squared column norms: warp w takes rows RPW w .. RPW w + RPW - 1 of the tile, lane l the columns l, l + 32, ...
        constexpr int RPW = IMP_TT / IMP_WARPS;
        const double* zn = zt + (tile % IMP_STAGES) * IMP_TT * KPT + RPW * warp * KPT;
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
          const int m = lane + 32 * q;
          if (m < k) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) nacc[q] = fma(zn[r * KPT + m], zn[r * KPT + m], nacc[q]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < IMP_MT; ++i) ImpactSlot<NT, NT>::run(nact[i], acc[i], zs + pa[i], zs + pb[i], ds);
    }
    cp_async_wait<0>();
    // scatter the fragments: thread holds pair (pa, pb) x equations j = 8 nt + 2 fc + {0, 1}
#pragma unroll
    for (int i = 0; i < IMP_MT; ++i) {
      const int mt = mt0 + warp + IMP_WARPS * i;
      if (mt >= mtiles || 8 * mt + fr >= npairs) continue;
      const int a = pa[i], b = pb[i];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt < NT - nact[i]) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 8 * nt + 2 * fc + h;
          if (j >= k) continue;
          if (j > a) G[impact_row_off(j) + tri_idx(a, b)] = acc[i][nt][h];                       // Gram of row j
          else if (j == a && b < a) G[impact_row_off(j) + j * (j + 1) / 2 + b] = acc[i][nt][h];  // rhs of row a
        }
      }
    }
  }
  // column norms: the eight warps' partial sums, added in warp order
  __syncthreads();
  double* part = sm;  // [IMP_WARPS][k], the staging buffers are free now
#pragma unroll
  for (int q = 0; q < NCOL; ++q) {
    const int m = lane + 32 * q;
    if (m < k) part[warp * k + m] = nacc[q];
  }
  __syncthreads();
  if (tid < k) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < IMP_WARPS; ++q) s += part[q * k + tid];
    P.znorm[(size_t)u * k + tid] = s;
  }
  if (pw == 0) return;  // the row draws run as their own batch-wide launches (kernels_impact_rows.cu)
  __threadfence_block();
  __syncthreads();
  impact_draw_rows<true>(D, P, G, wbuf, u, &bad, pw, 0);
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 2);
}

// ---- LDLT: the row draws from the unit's Z'Z (full k x kz) ---------------------------------------------------------
// mode bit 0: compute Z'Z here (small k: at most two passes of one pair per thread; the grouped contraction would spend a
// 128 x 128 tile on a 9 x 9 result), otherwise it was written by the gram_Z contraction; bit 1: Gram only, no draws.
__global__ void __launch_bounds__(256) impact_ldlt_kernel(const Dims D, const DevPtrs P, double* gscratch, const int gsize, const int kz, const int pw,
                                                          const int mode) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x;
  const int k = D.k;
  double* G = gscratch + (size_t)u * gsize;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  if (mode & 1) {
    constexpr int TT = 16;
    double* zt = sm;  // [TT][k] (the phase-2 workspaces start only after the Gram is done)
    const int n = P.win_n[w], ldM = D.ldM;
    const long long off = P.win_off[w] + (long long)c * k;
    const int npairs = k * (k + 1) / 2;
    for (int pbase = 0; pbase < npairs; pbase += blockDim.x) {
      const int pid = pbase + tid;
      int a = 0, b = 0;
      const bool active = pid < npairs;
      if (active) pair_of(pid, a, b);
      double acc = 0.0;
      for (int t0 = 0; t0 < n; t0 += TT) {
        __syncthreads();
        for (int e = tid; e < TT * k; e += blockDim.x) {
          const int tt = e / k, m = e % k;
          zt[e] = t0 + tt < n ? P.Z[off + (long long)(t0 + tt) * ldM + m] : 0.0;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
          for (int tt = 0; tt < TT; ++tt) acc = fma(zt[tt * k + a], zt[tt * k + b], acc);
        }
      }
      if (active) { G[a * kz + b] = acc; G[b * kz + a] = acc; }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (mode & 2) return;
  for (int a = tid; a < k; a += blockDim.x) P.znorm[(size_t)u * k + a] = G[a * kz + a];
  if (pw == 0) return;  // the row draws run as their own batch-wide launches (kernels_impact_rows.cu)
  __threadfence_block();
  __syncthreads();
  impact_draw_rows<false>(D, P, G, sm, u, &bad, pw, kz);
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 2);
}

// ---- LDLT updateState: d_i = 1 / Gamma(shape_i + floor(n/2), 1 / (scl_i + ||(Z L')_i||^2 / 2))  (draw.h:1244-1259) ------
// with ||(Z L')_i||^2 = L_i (Z'Z) L_i' from the Gram the contemporaneous draw already used: k^3/3 flops per unit instead of
// a second n k^2 / 2 pass over the residuals.  One CTA per unit, one warp per series; Z'Z staged in shared memory when it fits.
__global__ void __launch_bounds__(256) ldlt_state_kernel(const Dims D, const DevPtrs P, const double* gscratch, const int gsize, const int kz, const int in_smem) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = D.k, n = P.win_n[w];
  const double* G = gscratch + (size_t)u * gsize;
  if (in_smem) {
    for (int e = tid; e < k * kz; e += blockDim.x) sm[e] = G[e];
    __syncthreads();
    G = sm;
  }
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const double* Lu = P.L + (size_t)u * k * k;
  for (int i = warp; i < k; i += (blockDim.x >> 5)) {
    const double* Li = Lu + (size_t)i * k;
    double ss = 0.0;
    for (int a = lane; a <= i; a += 32) {  // t_a = sum_{b<=i} G_ab L_ib (G symmetric: column a read along rows b, coalesced in a)
      double t0 = 0.0, t1 = 0.0;
      int b = 0;
      for (; b + 1 <= i; b += 2) {
        t0 = fma(G[b * kz + a], Li[b], t0);
        t1 = fma(G[(b + 1) * kz + a], Li[b + 1], t1);
      }
      if (b <= i) t0 = fma(G[b * kz + a], Li[b], t0);
      ss = fma(Li[a], t0 + t1, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0)
      P.diag[(size_t)u * k + i] = 1.0 / rng.gamma(ST_LDLT_DIAG, (uint32_t)i, P.sig_shp[i] + (double)(n / 2), 1.0 / (P.sig_scl[i] + ss / 2));
  }
}

}  // namespace

bool ldlt_own_gram(const Dims& D) { return D.k < 24; }  // npairs <= 276: the draw kernel forms Z'Z itself

int impact_gsize(const Dims& D) {
  const int k = D.k;
  if (!D.sv) return k * ((k + 1) & ~1);  // full Z'Z, even row pitch
  return ((k - 1) * k * (2 * k - 1) / 6 + 3 * (k - 1) * k / 2) / 2 + 8;
}

// j-tile count of the kernel instance that serves k (see launch_impact)
static int impact_nt_bucket(int k) {
  const int nt = (k + 7) / 8;
  return nt <= 4 ? nt : nt <= 7 ? 7 : 13;
}
// shared memory with `pw` phase-2 workspaces (one per working warp)
static size_t impact_smem_pw(const Dims& D, int pw) {
  const int k = D.k;
  const size_t work = (size_t)pw * (k * (k + 1) / 2 + 3 * k);
  if (D.sv) return (2 * (size_t)IMP_STAGES * imp_tt(impact_nt_bucket(k)) * (8 * impact_nt_bucket(k) + 4) + ((k + 1) & ~1) + work) * sizeof(double);
  return std::max(work, (size_t)16 * k) * sizeof(double);
}
static int impact_warps(const Dims& D, size_t smem_limit) {  // as many of the 8 warps as fit a workspace
  int pw = IMP_WARPS;
  while (pw > 0 && impact_smem_pw(D, pw) > smem_limit) --pw;
  return pw;
}
size_t impact_smem(const Dims& D) { return impact_smem_pw(D, 1); }

// rows of at most 128 unknowns are drawn by the batch-wide warp-per-row kernels; wider models keep the in-CTA draws
static bool impact_split(const Dims& D) { return D.k - 1 <= 128 && getenv("BVHAR_IMPACT_FUSED") == nullptr; }

bool impact_supported(const Dims& D, size_t smem_limit) {
  return (impact_split(D) || impact_warps(D, smem_limit) >= 1) && (!D.sv || D.k <= 104);
}

// kernels one launch_impact call starts (the engine's launch counter)
int impact_launches(const Dims& D) {
  if (!impact_split(D)) return 1;
  const int jmax = D.k - 1;
  return 1 + (jmax > 64) + (jmax > 32) + 1;
}

cudaError_t launch_impact(const Dims& D, const DevPtrs& P, double* gscratch, cudaStream_t st) {
  const bool split = impact_split(D);
  const int pw = split ? 0 : impact_warps(D, 227 * 1024);
  if (!split && pw < 1) return cudaErrorInvalidConfiguration;
  const size_t smem = impact_smem_pw(D, pw);
  const int gs = impact_gsize(D);
  if (D.sv) {
#define BV_IMPACT(NTV, MTV)                                                                                   \
  do {                                                                                                        \
    cudaFuncSetAttribute(impact_sv_kernel<NTV, MTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    impact_sv_kernel<NTV, MTV><<<D.U, IMP_WARPS * 32, smem, st>>>(D, P, gscratch, gs, pw);                        \
  } while (0)
    const int nt = (D.k + 7) / 8;
    if (nt <= 1) BV_IMPACT(1, 4);
    else if (nt <= 2) BV_IMPACT(2, 4);
    else if (nt <= 3) BV_IMPACT(3, 4);
    else if (nt <= 4) BV_IMPACT(4, 4);
    else if (nt <= 7) BV_IMPACT(7, 3);
    else if (nt <= 13) BV_IMPACT(13, 2);
    else return cudaErrorInvalidConfiguration;
#undef BV_IMPACT
  } else {
    cudaFuncSetAttribute(impact_ldlt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    impact_ldlt_kernel<<<D.U, 256, smem, st>>>(D, P, gscratch, gs, (D.k + 1) & ~1, pw, ldlt_own_gram(D) ? 1 : 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !split) return e;
  const int kz = (D.k + 1) & ~1, jmax = D.k - 1;  // widest rows first
  if (jmax > 64) { e = launch_impact_rows_nc4(D, P, gscratch, gs, kz, 65, jmax, st); if (e != cudaSuccess) return e; }
  if (jmax > 32) { e = launch_impact_rows_nc2(D, P, gscratch, gs, kz, 33, std::min(jmax, 64), st); if (e != cudaSuccess) return e; }
  return launch_impact_rows_nc1(D, P, gscratch, gs, kz, 1, std::min(jmax, 32), st);
}

// Z'Z alone (stage-wise entry points: updateState of a small-k model run on its own)
cudaError_t launch_ldlt_gram_small(const Dims& D, const DevPtrs& P, double* gscratch, cudaStream_t st) {
  const size_t smem = impact_smem_pw(D, 1);
  cudaFuncSetAttribute(impact_ldlt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  impact_ldlt_kernel<<<D.U, 256, smem, st>>>(D, P, gscratch, impact_gsize(D), (D.k + 1) & ~1, 1, 3);
  return cudaGetLastError();
}
cudaError_t launch_ldlt_state(const Dims& D, const DevPtrs& P, const double* gscratch, cudaStream_t st) {
  const int kz = (D.k + 1) & ~1;
  const size_t need = (size_t)D.k * kz * sizeof(double);
  const int in_smem = need <= 200 * 1024 ? 1 : 0;
  const size_t smem = in_smem ? need : 0;
  if (in_smem) cudaFuncSetAttribute(ldlt_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ldlt_state_kernel<<<D.U, 256, smem, st>>>(D, P, gscratch, impact_gsize(D), kz, in_smem);
  return cudaGetLastError();
}

}  // namespace bvhar
