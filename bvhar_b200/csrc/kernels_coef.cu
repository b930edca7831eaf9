// kernels_coef.cu — stage 4 (updateCoef, triangular.h:281-316 + draw_coef draw.h:372-383 + draw_mn_savs
// draw.h:417-430) split by data dependence into three kernels:
//
//   coef_pbuild   (SV)  P_j = sum_{i>=j} L_ij^2 S_i for all j of a unit in ONE pass over S (each S_i read once)
//   coef_chol           chol(diag(prec_j) + P_j) for every (unit, equation): no dependence on this sweep's
//                       draws, so all U*k factorisations run concurrently, one warp per matrix
//   coef_seq            the part that IS sequential in j (equation j conditions on the draws of equations < j):
//                       with g_i = S_i v_i, v_i = A L_i' kept up to date,
//                          r_j     = prec_j o (mean_j - a_j) + sum_{i>=j} L_ij (C_i - g_i)
//                          delta_j = P_j^-1 r_j + chol(P_j)^-T z,   a_j += delta_j,   g_i += L_ij S_i delta_j  (i > j)
//                       which is the reference's draw (a_j = P_j^-1 b_j + chol^-T z with equation j zeroed,
//                       tri.h:283) written as an increment, so no product with the old a_j is ever needed.
//
// Packed symmetric products S_i x are done by one warp straight from global/L2 memory, row by row with
// coalesced loads: lanes own columns (private accumulators for the upper part) and one warp reduction
// per row gives the lower part.  LDLT models have S_i = X'X / d_i: one product per equation.
#include "coef_sizes.h"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "warp_linalg.cuh"

#ifndef BV_NC
#error "compile with -DBV_NC=1|2|4|8 (one translation unit per ceil(d / 32) class)"
#endif
#define BV_CAT2(a, b) a##b
#define BV_CAT(a, b) BV_CAT2(a, b)

namespace bvhar {

namespace {

// ---- B: the sequential-in-j part, one CTA per unit ---------------------------------------------
// Shared memory holds everything the dependent chain touches: h_i = C_i - S_i v_i (k x d), the unit's
// coefficient matrix, all k*d standard normals of the sweep (drawn up front: they depend on nothing), and the
// Cholesky factor of the current equation, double-buffered so that the factor of equation j + 1 streams in
// (cp.async) while equation j is solved.  Per equation only the solve (one warp) and the k-j-1 products
// S_i delta (all warps, S_i from L2) remain on the critical path.
template <bool SV, int NC>
__global__ void __launch_bounds__(SEQ_THREADS, (NC <= 2 ? 4 : 1)) coef_seq_kernel(const Dims D, const DevPtrs P, const double* __restrict__ Pfac) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int k = D.k, d = D.d, ldM = D.ldM;
  const int ppad = D.ldS;                   // even pitch of a packed factor (16-byte chunks)
  double* Lc = sm;                          // [2][ppad] factor of the current / next equation
  double* h = Lc + 2 * (size_t)ppad;        // [k][d]  C_i - S_i v_i
  double* Asm = h + (size_t)k * d;          // [k][d]  coefficient matrix, equation-major
  double* zall = Asm + (size_t)k * d;       // [k][d]  this sweep's normals
  double* Lsm = zall + (size_t)k * d;       // [k][k]
  double* rv = Lsm + k * k;                 // [d]
  double* dl = rv + d;                      // delta
  double* tv = dl + d;                      // LDLT: X'X delta
  double* wscr = tv + 2 * d;                // [nw][d] per-warp vectors
  double* idg = wscr + (size_t)nw * d;      // [k] 1 / diag (LDLT)
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  double* Ac = P.Acoef + (size_t)w * d * ldM + (size_t)c * k;
  const double* prec_u = P.prec + (size_t)u * k * d;
  const double* xn = P.xnorm + (size_t)w * d;
  const double* X2 = SV ? nullptr : P.XtX + (size_t)w * D.ldS;

  auto prefetch_factor = [&](int j) {
    const double* Lg = Pfac + ((size_t)u * k + j) * D.ldS;
    double* dst = Lc + (size_t)(j & 1) * ppad;
    for (int p = 2 * tid; p < ppad; p += 2 * SEQ_THREADS) cp_async16(dst + p, Lg + p, 16);
    cp_async_commit();
  };
  prefetch_factor(0);

  for (int e = tid; e < k * k; e += blockDim.x) Lsm[e] = P.L[(size_t)u * k * k + e];
  for (int e = tid; e < d * k; e += blockDim.x) {
    const int a = e / k, m = e - a * k;
    Asm[m * d + a] = Ac[(size_t)a * ldM + m];
  }
  for (int e = tid; e < (k * d + 1) / 2; e += blockDim.x) {  // ST_COEF_Z element j*d + a, pairs (2e, 2e+1)
    double z0, z1;
    rng.normal_pair(ST_COEF_Z, (uint32_t)e, 0, z0, z1);
    zall[2 * e] = z0;
    if (2 * e + 1 < k * d) zall[2 * e + 1] = z1;
  }
  if (!SV) for (int i = tid; i < k; i += blockDim.x) idg[i] = 1.0 / P.diag[(size_t)u * k + i];
  __syncthreads();
  // h_i = C_i - S_i v_i, v_i = A L_i'
  for (int i = warp; i < k; i += nw) {
    double* v = wscr + (size_t)warp * d;
    for (int a = lane; a < d; a += 32) {
      double sacc = 0.0;
      for (int m = 0; m <= i; ++m) sacc = fma(Lsm[i * k + m], Asm[m * d + a], sacc);
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

  for (int j = 0; j < k; ++j) {
    cp_async_wait<0>();
    __syncthreads();  // factor j landed; h, Asm of the previous equation are final
    if (j + 1 < k) prefetch_factor(j + 1);
    const double* Lj = Lc + (size_t)(j & 1) * ppad;
    for (int a = tid; a < d; a += blockDim.x) {
      double acc = prec_u[j * d + a] * (P.prior_mean[j * d + a] - Asm[j * d + a]);
      for (int i = j; i < k; ++i) acc = fma(Lsm[i * k + j], h[(size_t)i * d + a], acc);
      rv[a] = acc;
    }
    __syncthreads();
    if (warp == 0) precision_solve_blocked<NC>(Lj, d, rv, zall + (size_t)j * d, dl);
    __syncthreads();
    for (int a = tid; a < d; a += blockDim.x) {
      const double val = Asm[j * d + a] + dl[a];
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
    if (j + 1 < k) {
      if (SV) {
        for (int i = j + 1 + warp; i < k; i += nw) {
          double out[NC];
          symv_packed_warp<NC>(P.S + ((size_t)u * k + i) * D.ldS, dl, d, out);
          const double lij = Lsm[i * k + j];
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) h[(size_t)i * d + b] = fma(-lij, out[sidx], h[(size_t)i * d + b]);
          }
        }
      } else {
        if (warp == 0) {
          double out[NC];
          symv_packed_warp<NC>(X2, dl, d, out);
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) tv[b] = out[sidx];
          }
        }
        __syncthreads();
        for (int e = tid; e < (k - j - 1) * d; e += blockDim.x) {
          const int i = j + 1 + e / d, a = e % d;
          h[(size_t)i * d + a] = fma(-Lsm[i * k + j] * idg[i], tv[a], h[(size_t)i * d + a]);
        }
      }
    }
  }
}

// ---- B' (SV): the same sequential part, software-pipelined across equations ------------------------------------
// Only Q_j delta_j, with Q_j = sum_{i>=j+1} L_{i,j+1} L_ij S_i precomputed by coef_pbuild_kernel, separates the solve of
// equation j from the solve of equation j + 1:
//     r_{j+1} = rbase_{j+1} - Q_j delta_j,   rbase_{j+1} = prec o (mean - a_{j+1}) + sum_{i>=j+1} L_{i,j+1} h_i^{(lag)}
// where h^{(lag)} carries the updates of delta_0 .. delta_{j-1} only.  So while warp 0 solves equation j, warps 1..7
// apply delta_{j-1} to the h_i (k-j-1 products S_i delta from L2), build rbase_{j+1} and prefetch the next factor and the
// next Q; the critical path per equation is one 15 KB shared-memory matvec plus the solve.
constexpr int PIPE_WORKERS = SEQ_THREADS - 32;  // threads of warps 1..7
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(PIPE_WORKERS)); }

template <int NC>
__global__ void __launch_bounds__(SEQ_THREADS) coef_seq_pipe_kernel(const Dims D, const DevPtrs P, const double* __restrict__ Pfac,
                                                                     const double* __restrict__ Qn) {
  extern __shared__ __align__(16) double sm[];
  const int u = blockIdx.x, w = u / D.C, c = u % D.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int k = D.k, d = D.d, ldM = D.ldM;
  const int ppad = D.ldS;
  double* Lc = sm;                          // [2][ppad] Cholesky factor of equation j / j + 1
  double* Qc = Lc + 2 * (size_t)ppad;       // [2][ppad] Q_{j-1} / Q_j
  double* h = Qc + 2 * (size_t)ppad;        // [k][d]
  double* Asm = h + (size_t)k * d;          // [k][d]
  double* zall = Asm + (size_t)k * d;       // [k][d]
  double* Lsm = zall + (size_t)k * d;       // [k][k]
  double* rv = Lsm + k * k;                 // [d] right-hand side of the current equation
  double* rbase = rv + d;                   // [2][d]
  double* dl = rbase + 2 * d;               // [2][d] delta_j (even / odd equations)
  double* wscr = dl + 2 * d + d;            // [nw][d]
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  double* Ac = P.Acoef + (size_t)w * d * ldM + (size_t)c * k;
  const double* prec_u = P.prec + (size_t)u * k * d;
  const double* xn = P.xnorm + (size_t)w * d;

  auto prefetch = [&](const double* src, double* dst, int t0, int nthr) {
    for (int p = 2 * t0; p < ppad; p += 2 * nthr) cp_async16(dst + p, src + p, 16);
  };
  prefetch(Pfac + ((size_t)u * k) * D.ldS, Lc, tid, SEQ_THREADS);
  cp_async_commit();

  for (int e = tid; e < k * k; e += blockDim.x) Lsm[e] = P.L[(size_t)u * k * k + e];
  for (int e = tid; e < d * k; e += blockDim.x) {
    const int a = e / k, m = e - a * k;
    Asm[m * d + a] = Ac[(size_t)a * ldM + m];
  }
  for (int e = tid; e < (k * d + 1) / 2; e += blockDim.x) {
    double z0, z1;
    rng.normal_pair(ST_COEF_Z, (uint32_t)e, 0, z0, z1);
    zall[2 * e] = z0;
    if (2 * e + 1 < k * d) zall[2 * e + 1] = z1;
  }
  __syncthreads();
  for (int i = warp; i < k; i += nw) {  // h_i = C_i - S_i v_i, v_i = A L_i'
    double* v = wscr + (size_t)warp * d;
    for (int a = lane; a < d; a += 32) {
      double sacc = 0.0;
      for (int m = 0; m <= i; ++m) sacc = fma(Lsm[i * k + m], Asm[m * d + a], sacc);
      v[a] = sacc;
    }
    __syncwarp();
    double out[NC];
    symv_packed_warp<NC>(P.S + ((size_t)u * k + i) * D.ldS, v, d, out);
#pragma unroll
    for (int sidx = 0; sidx < NC; ++sidx) {
      const int b = lane + 32 * sidx;
      if (b < d) h[(size_t)i * d + b] = P.Cm[((size_t)u * k + i) * D.ldX + b] - out[sidx];
    }
    __syncwarp();
  }
  __syncthreads();
  for (int a = tid; a < d; a += blockDim.x) {  // rbase_0
    double acc = prec_u[a] * (P.prior_mean[a] - Asm[a]);
    for (int i = 0; i < k; ++i) acc = fma(Lsm[i * k], h[(size_t)i * d + a], acc);
    rbase[a] = acc;
  }
  // threads-per-row of the shared-memory matvec (power of two, tpr * d <= 256)
  int tpr = 1;
  while (tpr < 8 && 2 * tpr * d <= SEQ_THREADS) tpr <<= 1;

  for (int j = 0; j < k; ++j) {
    cp_async_wait<0>();
    __syncthreads();  // factor j and Q_{j-1} landed; delta_{j-1}, rbase_j, Asm of the previous equation are final
    // ---- phase A (all warps): r_j = rbase_j - Q_{j-1} delta_{j-1}
    {
      const double* Qp = Qc + (size_t)((j - 1) & 1) * ppad;
      const double* dprev = dl + (size_t)((j - 1) & 1) * d;
      const double* rb = rbase + (size_t)(j & 1) * d;
      const int a = tid / tpr, sub = tid - a * tpr;
      double acc = 0.0;
      if (j > 0 && a < d) {
        const int ra = a * (a + 1) / 2;
        for (int b = sub; b < d; b += tpr) acc = fma(b <= a ? Qp[ra + b] : Qp[b * (b + 1) / 2 + a], dprev[b], acc);
      }
      for (int o = tpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (sub == 0 && a < d) rv[a] = rb[a] - acc;
    }
    __syncthreads();
    // ---- phase B
    double* dcur = dl + (size_t)(j & 1) * d;
    if (warp == 0) {
      precision_solve_blocked<NC>(Lc + (size_t)(j & 1) * ppad, d, rv, zall + (size_t)j * d, dcur);
    } else {
      const int wt = tid - 32;  // 0 .. PIPE_WORKERS-1
      if (j + 1 < k) {
        prefetch(Pfac + ((size_t)u * k + j + 1) * D.ldS, Lc + (size_t)((j + 1) & 1) * ppad, wt, PIPE_WORKERS);
        prefetch(Qn + ((size_t)u * k + j) * D.ldS, Qc + (size_t)(j & 1) * ppad, wt, PIPE_WORKERS);
      }
      cp_async_commit();
      if (j > 0) {  // h_i -= L_{i,j-1} S_i delta_{j-1}, i >= j + 1
        const double* dprev = dl + (size_t)((j - 1) & 1) * d;
        for (int i = j + 1 + (warp - 1); i < k; i += nw - 1) {
          double out[NC];
          symv_packed_warp<NC>(P.S + ((size_t)u * k + i) * D.ldS, dprev, d, out);
          const double lij = Lsm[i * k + j - 1];
#pragma unroll
          for (int sidx = 0; sidx < NC; ++sidx) {
            const int b = lane + 32 * sidx;
            if (b < d) h[(size_t)i * d + b] = fma(-lij, out[sidx], h[(size_t)i * d + b]);
          }
        }
      }
      worker_barrier();
      if (j + 1 < k) {  // rbase_{j+1} from h^{(lag)} (delta_j not applied: it enters through Q_j in the next phase A)
        double* rb = rbase + (size_t)((j + 1) & 1) * d;
        for (int a = wt; a < d; a += PIPE_WORKERS) {
          double acc = prec_u[(j + 1) * d + a] * (P.prior_mean[(j + 1) * d + a] - Asm[(j + 1) * d + a]);
          for (int i = j + 1; i < k; ++i) acc = fma(Lsm[i * k + j + 1], h[(size_t)i * d + a], acc);
          rb[a] = acc;
        }
      }
    }
    __syncthreads();
    for (int a = tid; a < d; a += blockDim.x) {
      const double val = Asm[j * d + a] + dcur[a];
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
  }
}

}  // namespace

// part 2 of the coefficient stage for designs with ceil(d / 32) <= BV_NC
cudaError_t BV_CAT(launch_coef_seq_nc, BV_NC)(const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st) {
  constexpr int NC = BV_NC;
  const int k = D.k, d = D.d;
  if (D.sv && Qn) {
    const size_t smp = coef_pipe_smem_doubles(k, d, D.ldS) * sizeof(double);
    if (smp > smem_limit) return cudaErrorInvalidConfiguration;
    cudaFuncSetAttribute(coef_seq_pipe_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smp);
    coef_seq_pipe_kernel<NC><<<D.U, SEQ_THREADS, smp, st>>>(D, P, Pfac, Qn);
    return cudaGetLastError();
  }
  const size_t sm3 = coef_seq_smem_doubles<true>(k, d, D.ldS) * sizeof(double);
  if (sm3 > smem_limit) return cudaErrorInvalidConfiguration;
  if (D.sv) {
    cudaFuncSetAttribute(coef_seq_kernel<true, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
    coef_seq_kernel<true, NC><<<D.U, SEQ_THREADS, sm3, st>>>(D, P, Pfac);
  } else {
    cudaFuncSetAttribute(coef_seq_kernel<false, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
    coef_seq_kernel<false, NC><<<D.U, SEQ_THREADS, sm3, st>>>(D, P, Pfac);
  }
  return cudaGetLastError();
}

}  // namespace bvhar
