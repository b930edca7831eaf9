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
//            j >= a are issued.  LDLT models have w constant in t: one plain Z'Z per unit (FMA).
//   phase 2  one warp per row: packed Cholesky of the j x j precision, the two triangular solves, the draw,
//            the SAVS copy, and the rebuilt row of L.
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

constexpr int IMP_TT = 32;      // time points per staged tile
constexpr int IMP_STAGES = 3;   // cp.async pipeline depth
constexpr int IMP_WARPS = 8;

// ---- phase 2, shared by both covariance models ---------------------------------------------------------
template <bool SV>
__device__ void impact_draw_rows(const Dims& D, const DevPtrs& P, const double* G, double* wbuf, int u, int* bad, int nw) {
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
    for (int p = lane; p < j * (j + 1) / 2; p += 32) Pw[p] = Gj[p] * inv_dj;
    __syncwarp();
    for (int b = lane; b < j; b += 32) {
      const double pr = P.chol_prec[(size_t)u * q + id + b];
      Pw[tri_idx(b, b)] += pr;
      r[b] = SV ? Gj[j * (j + 1) / 2 + b] : G[tri_idx(j, b)] * inv_dj;  // prior_chol_mean = 0 (tri.h:52)
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
template <int NT, int IMP_MT>
__global__ void __launch_bounds__(IMP_WARPS * 32) impact_sv_kernel(const Dims D, const DevPtrs P, double* gscratch, const int gsize, const int pw) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int k = D.k, ldM = D.ldM;
  const int n = P.win_n[w];
  const long long off = P.win_off[w] + (long long)c * k;
  const int KP = k | 1;  // odd row pitch
  double* zt = sm;                                   // [STAGES][TT][KP]
  double* dt = zt + IMP_STAGES * IMP_TT * KP;        // [STAGES][TT][KP]
  double* nrm = dt + IMP_STAGES * IMP_TT * KP;       // [k] squared column norms of Z
  double* wbuf = nrm + ((k + 1) & ~1);               // phase-2 per-warp workspaces
  double* G = gscratch + (size_t)u * gsize;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  const int npairs = k * (k + 1) / 2;
  const int mtiles = (npairs + 7) / 8;
  const int ntiles = (n + IMP_TT - 1) / IMP_TT;
  const int per_pass = IMP_MT * IMP_WARPS;
  double nacc = 0.0;  // thread i < k: sum_t z_ti^2 (first pass only)

  auto stage_load = [&](int tile, int slot) {
    const int t0 = tile * IMP_TT;
    double* zs = zt + slot * IMP_TT * KP;
    double* ds = dt + slot * IMP_TT * KP;
    for (int e = tid; e < IMP_TT * k; e += IMP_WARPS * 32) {
      const int tt = e / k, m = e - tt * k;
      const bool ok = t0 + tt < n;
      const long long at = ok ? off + (long long)(t0 + tt) * ldM + m : off;
      cp_async8(zs + tt * KP + m, P.Z + at, ok ? 8 : 0);
      cp_async8(ds + tt * KP + m, P.Dinv + at, ok ? 8 : 0);
    }
  };

  for (int mt0 = 0; mt0 < mtiles; mt0 += per_pass) {
    // this warp's m-tiles of the pass: mt0 + warp + 8 i (heavy early tiles and light late tiles interleave)
    int pa[IMP_MT], pb[IMP_MT], ntlo[IMP_MT];
    bool live[IMP_MT];
#pragma unroll
    for (int i = 0; i < IMP_MT; ++i) {
      const int mt = mt0 + warp + IMP_WARPS * i;
      live[i] = mt < mtiles;
      int a0, b0;
      pair_of(min(8 * mt, npairs - 1), a0, b0);
      ntlo[i] = a0 >> 3;  // tiles of j below the smallest a of the m-tile are never needed (j >= a)
      pair_of(min(8 * mt + fr, npairs - 1), pa[i], pb[i]);
    }
    double acc[IMP_MT][NT][2];
#pragma unroll
    for (int i = 0; i < IMP_MT; ++i)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;

    __syncthreads();  // the previous pass is done with the staging buffers
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
      const double* zs = zt + (tile % IMP_STAGES) * IMP_TT * KP;
      const double* ds = dt + (tile % IMP_STAGES) * IMP_TT * KP;
      if (mt0 == 0 && tid < k) {
#pragma unroll 8
        for (int tt = 0; tt < IMP_TT; ++tt) nacc = fma(zs[tt * KP + tid], zs[tt * KP + tid], nacc);
      }
#pragma unroll 2
      for (int ks = 0; ks < IMP_TT; ks += 4) {
        const double* zr = zs + (ks + fc) * KP;
        const double* dr = ds + (ks + fc) * KP;
        double bf[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bf[nt] = dr[min(8 * nt + fr, k - 1)];
#pragma unroll
        for (int i = 0; i < IMP_MT; ++i) {
          if (!live[i]) continue;  // warp-uniform
          const double af = zr[pa[i]] * zr[pb[i]];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            if (nt >= ntlo[i]) dmma_m8n8k4(acc[i][nt][0], acc[i][nt][1], af, bf[nt]);
        }
      }
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
        if (nt < ntlo[i]) continue;
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
  if (tid < k) P.znorm[(size_t)u * k + tid] = nacc;
  __threadfence_block();
  __syncthreads();
  impact_draw_rows<true>(D, P, G, wbuf, u, &bad, pw);
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 2);
}

// ---- LDLT: one Z'Z per unit (weights are constant in time), then the row draws ---------------------------
__global__ void __launch_bounds__(256) impact_ldlt_kernel(const Dims D, const DevPtrs P, double* gscratch, const int gsize, const int pw) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x;
  const int k = D.k, ldM = D.ldM;
  const int n = P.win_n[w];
  const long long off = P.win_off[w] + (long long)c * k;
  constexpr int TT = 16;
  double* zt = sm;            // [TT][k]
  double* wbuf = zt + TT * k; // per-warp packed matrices
  double* G = gscratch + (size_t)u * gsize;
  __shared__ int bad;
  if (tid == 0) bad = 0;
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
    if (active) {
      G[tri_idx(a, b)] = acc;
      if (a == b) P.znorm[(size_t)u * k + a] = acc;
    }
  }
  __threadfence_block();
  __syncthreads();
  impact_draw_rows<false>(D, P, G, wbuf, u, &bad, pw);
  __syncthreads();
  if (tid == 0 && bad) atomicOr(&P.flags[u], 2);
}

}  // namespace

int impact_gsize(const Dims& D) {
  const int k = D.k;
  if (!D.sv) return k * (k + 1) / 2;
  return ((k - 1) * k * (2 * k - 1) / 6 + 3 * (k - 1) * k / 2) / 2 + 8;
}

// shared memory with `pw` phase-2 workspaces (one per working warp)
static size_t impact_smem_pw(const Dims& D, int pw) {
  const int k = D.k;
  const size_t work = (size_t)pw * (k * (k + 1) / 2 + 3 * k);
  if (D.sv) return (2 * (size_t)IMP_STAGES * IMP_TT * (k | 1) + ((k + 1) & ~1) + work) * sizeof(double);
  return (16 * (size_t)k + work) * sizeof(double);
}
static int impact_warps(const Dims& D, size_t smem_limit) {  // as many of the 8 warps as fit a workspace
  int pw = IMP_WARPS;
  while (pw > 0 && impact_smem_pw(D, pw) > smem_limit) --pw;
  return pw;
}
size_t impact_smem(const Dims& D) { return impact_smem_pw(D, 1); }

bool impact_supported(const Dims& D, size_t smem_limit) {
  return impact_warps(D, smem_limit) >= 1 && (!D.sv || D.k <= 104);
}

cudaError_t launch_impact(const Dims& D, const DevPtrs& P, double* gscratch, cudaStream_t st) {
  const int pw = impact_warps(D, 227 * 1024);
  if (pw < 1) return cudaErrorInvalidConfiguration;
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
    impact_ldlt_kernel<<<D.U, 256, smem, st>>>(D, P, gscratch, gs, pw);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
