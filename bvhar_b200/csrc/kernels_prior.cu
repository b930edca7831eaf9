// kernels_prior.cu — shrinkage-prior hyper-parameter updaters as fused per-unit kernels
// (updateCoefPrec / updateImpactPrec of the seven prior classes,
//  inst/include/bvhar/src/bayes/triangular/triangular.h:455-457, 513-531, 599-623, 693-720, 793-818,
//  885-908, 974-993; draw primitives bayes/misc/draw.h:269-361, 649-749, 758-890, 956-998,
//  1007-1139, 1147-1235).
// One CTA per unit; the N (or q) elementwise conditional draws are strided over the threads, the
// group / global reductions are deterministic block reductions, griddy-Gibbs steps evaluate the grid
// in parallel and draw the index with one thread.  Draw sites use Philox stages
// ST_PRIOR_COEF + s (coefficients) and ST_PRIOR_CONTEM + s (contemporaneous terms), s = order of
// appearance in the reference's updater (DESIGN.md "Random streams").
// Reference quirks kept on purpose are marked QUIRK.
#include "engine.cuh"
#include "kernels.h"
#include "rng.cuh"

namespace bvhar {

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
struct Blk {
  double* red;   // >= 40 doubles
  double* grid;  // >= max grid doubles
  int* pick;
  __device__ double sum(double v) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = wsum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double t = lane < nw ? red[lane] : 0.0;
      t = wsum(t);
      if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
  }
  // Index drawn from exp(logw - max) (common.h:577-580); grid[] holds logw.  NaN weights count as 0.
  __device__ int griddy_pick(const Philox& rng, uint32_t stage, int n) const {
    __syncthreads();
    if (threadIdx.x == 0) {
      double mx = -1.0 / 0.0;
      for (int i = 0; i < n; ++i) {
        const double v = grid[i];
        if (v == v && v > mx) mx = v;
      }
      double tot = 0.0;
      for (int i = 0; i < n; ++i) {
        double e = exp(grid[i] - mx);
        if (!(e == e)) e = 0.0;
        grid[i] = e;
        tot += e;
      }
      const double target = rng.uniform(stage, 0, 0) * tot;
      double cum = 0.0;
      int id = n - 1;
      for (int i = 0; i < n; ++i) {
        cum += grid[i];
        if (cum > target) { id = i; break; }
      }
      *pick = id;
    }
    __syncthreads();
    return *pick;
  }
};
__device__ __forceinline__ double interior_grid(int grid, int i) { return (double)(i + 1) * (1.0 / (double)(grid + 1)); }

// Writes prior_alpha_prec for alpha element e (alpha order) into the internal (j*d + a) layout.
__device__ __forceinline__ void put_prec(const Dims& D, double* prec_u, int e, double v) {
  prec_u[(e / D.nrow) * D.d + (e % D.nrow)] = v;
}

// horseshoe_global_sparsity, draw.h:666-677: sqrt(1 / Gamma(floor((m+1)/2), 1 / (1/latent + s)))
__device__ __forceinline__ double hs_global(const Philox& rng, uint32_t st, uint32_t idx, int m, double latent, double s) {
  return sqrt(1.0 / rng.gamma(st, idx, (double)((m + 1) / 2), 1.0 / (1.0 / latent + s)));
}
// dl_dir_griddy grid, draw.h:881: LinSpaced(grid, 1 / size, .5) with INTEGER 1 / size  (QUIRK)
__device__ __forceinline__ double dl_grid_point(int grid, int nc, int i) {
  const double lo0 = (double)(1 / nc);
  const double lo = lo0 < .5 ? lo0 : .5, hi = lo0 < .5 ? .5 : lo0;
  const double step = grid > 1 ? (hi - lo) / (grid - 1) : 0.0;
  return i == grid - 1 ? hi : lo + i * step;
}
// ng_shape_jump, draw.h:1098-1119
__device__ double ng_shape_jump(const Philox& rng, uint32_t st_z, uint32_t st_u, uint32_t idx, double gam, int nc,
                                double slog, double ssq, double glob, double sd) {
  const double cand = exp(log(gam) + rng.normal(st_z, idx) * sd);
  double lr = log(cand) - log(gam) + nc * (lgamma(gam) - lgamma(cand));
  lr += nc * cand * (log(cand) - 2 * log(glob));
  lr -= nc * gam * (log(gam) - 2 * log(glob));
  lr += (cand - gam) * slog;
  lr += (gam - cand) * ssq / (glob * glob);
  if (log(rng.uniform(st_u, idx)) < fmin(lr, 0.0)) return cand;
  return gam;
}

// sum_e log1p(x_e r) over e = start, start + stride, ... as the logarithm of a running product, renormalised every 8
// factors: one FMA and one multiply per element instead of a double-precision log1p (the GDP griddy steps evaluate
// grid x N of them per unit and sweep: 3 M at k = 100, d = 301).  Semantics of the sum of log1p kept: NaN when a factor is
// negative (log1p of an argument below -1), -inf when one is zero.  Absolute error <= N ulp(1), far below the
// O(N) magnitude of the grid's log-weights.
__device__ __forceinline__ double log1p_sum(const double* __restrict__ x, int n, int start, int stride, double r) {
  double p = 1.0;
  int ex = 0, cnt = 0;
  bool bad = false;
  for (int e = start; e < n; e += stride) {
    const double f = fma(x[e], r, 1.0);
    bad |= !(f >= 0.0);
    p *= f;
    if (++cnt == 8) {
      int qe;
      p = frexp(p, &qe);
      ex += qe;
      cnt = 0;
    }
  }
  if (bad) return 0.0 / 0.0;
  return log(p) + (double)ex * 0.6931471805599453094;
}

}  // namespace

// MAXT = 512: the 16-warp launch of long coefficient vectors; MAXT = 256: four CTAs per SM (64 registers; measured 2 / 3 /
// 4 / 6 CTAs: 0.36 / 0.31 / 0.29 / 0.30 ms at C3 - the kernel is latency-bound: one
// CTA per unit, serial reductions and rejection samplers)
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) prior_kernel(const Dims D, const DevPtrs P, const PriorConst C, const int side) {
  __shared__ double red[40];
  __shared__ double gridbuf[512];
  __shared__ double gsm[3][64];  // per-group scratch (G <= 64)
  __shared__ int pick;
  const int u = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int N = D.N, q = D.q, G = D.G;
  const Philox rng(P.seeds[u], (uint32_t)(D.ubase + u), *P.iter);
  const Blk B{red, gridbuf, &pick};
  double* hs = P.hs + (size_t)u * HS_SLOTS;
  const uint32_t P0 = ST_PRIOR_COEF, P1 = ST_PRIOR_CONTEM;

  if (side == 0) {
    const double* coef = P.alpha + (size_t)u * N;
    double* prec_u = P.prec + (size_t)u * D.k * D.d;
    double* hN0 = P.hN0 + (size_t)u * N;
    double* hN1 = P.hN1 + (size_t)u * N;
    double* hN2 = P.hN2 + (size_t)u * N;
    double* hG0 = P.hG0 + (size_t)u * G;
    double* hG1 = P.hG1 + (size_t)u * G;
    const int* gi = P.grp_idx;
    switch (D.prior) {
      case 1: break;  // Minnesota: fixed prior (tri.h:455)
      case 4: {       // hierarchical Minnesota, tri.h:513-524 + minnesota_lambda draw.h:956-965
        const double lam = hs[HS_OWN_LAMBDA];
        double chi = 0.0;
        for (int e = tid; e < N; e += nt) {
          const int at = (e / D.nrow) * D.d + (e % D.nrow);
          const double pr = prec_u[at] * lam;
          const double df = coef[e] - P.prior_mean[at];
          chi += df * df * pr;
        }
        chi = B.sum(chi);
        __shared__ double newlam;
        if (tid == 0) {
          newlam = cut_param(rng.gig(P0 + 0, 0, C.hmn_shape - (double)(N / 2), 2 * C.hmn_rate, chi));
          hs[HS_OWN_LAMBDA] = newlam;
        }
        __syncthreads();
        // QUIRK: the cross-lag scale nu is drawn on a grid but touches no coefficient, because
        // RegParams::_cross_id stays empty (set_grp_id takes it by value, draw.h:75); not drawn here.
        for (int e = tid; e < N; e += nt) {
          const int at = (e / D.nrow) * D.d + (e % D.nrow);
          prec_u[at] = prec_u[at] * lam / newlam;
        }
        break;
      }
      case 2: {  // SSVS, tri.h:599-616
        double* dummy = hN0;
        double* slab = hN1;
        double* weight = hG0;
        const double spike = hs[HS_SPIKE];
        double ss = 0.0;
        for (int e = tid; e < N; e += nt) {  // ssvs_local_slab draw.h:336-345
          const double cf = coef[e];
          const double s = sqrt(1.0 / rng.gamma(P0 + 0, (uint32_t)e, C.ssvs_coef_slab_shape + .5,
                                                1.0 / (C.ssvs_coef_slab_scl + cf * cf / (dummy[e] + (1 - dummy[e]) * spike))));
          slab[e] = s;
          const double r = cf / s;
          ss += r * r;
        }
        ss = B.sum(ss);
        const int grid = C.ssvs_coef_grid;  // ssvs_scl_griddy draw.h:347-361
        for (int i = tid; i < grid; i += nt) {
          const double c = interior_grid(grid, i);
          gridbuf[i] = -ss / (2 * c * c) - N * log(c);
        }
        const double nspike = interior_grid(grid, B.griddy_pick(rng, P0 + 1, grid));
        for (int e = tid; e < N; e += nt) {  // ssvs_dummy draw.h:269-281
          const double cf = coef[e], sn = slab[e], sdn = nspike * slab[e], wt = weight[gi[e]];
          double e1 = -cf * cf / (2 * sn * sn), e2 = -cf * cf / (2 * sdn * sdn);
          const double mx = fmax(e1, e2);
          e1 = wt * exp(e1 - mx) / sn;
          e2 = (1 - wt) * exp(e2 - mx) / sdn;
          const double dm = rng.bernoulli(P0 + 2, (uint32_t)e, e1 / (e1 + e2));
          dummy[e] = dm;
          put_prec(D, prec_u, e, 1.0 / (nspike * (1 - dm) * sn + dm * sn));  // tri.h:615 (first power)  QUIRK
        }
        __syncthreads();
        for (int g = 0; g < G; ++g) {  // ssvs_mn_weight draw.h:308-325
          double sm = 0.0;
          for (int e = tid; e < N; e += nt)
            if (gi[e] == g) sm += dummy[e];
          sm = B.sum(sm);
          if (tid == 0) weight[g] = rng.beta(P0 + 3, (uint32_t)g, P.ssvs_s1[g] + sm, P.ssvs_s2[g] + P.grp_cnt[g] - sm);
        }
        if (tid == 0) hs[HS_SPIKE] = nspike;
        break;
      }
      case 3: {  // Horseshoe, tri.h:693-711
        double* local = hN0;
        double* latent_local = hN1;
        double* shrink = hN2;
        double* group = hG0;
        double* latent_group = hG1;
        const double glob_old = D.ggl ? hs[HS_GLOBAL] : 1.0;
        for (int g = 0; g < G; ++g) {
          double s = 0.0;
          for (int e = tid; e < N; e += nt)
            if (gi[e] == g) {
              const double ml = glob_old * local[e];
              s += coef[e] * coef[e] / (2 * ml * ml);
            }
          s = B.sum(s);
          if (tid == 0) {
            const double lg = 1.0 / rng.gamma(P0 + 0, (uint32_t)g, 1.0, 1.0 / (1 + 1 / (group[g] * group[g])));  // horseshoe_latent
            latent_group[g] = lg;
            gsm[0][g] = hs_global(rng, P0 + 1, (uint32_t)g, P.grp_cnt[g], lg, s);  // horseshoe_mn_sparsity draw.h:713-733
            group[g] = gsm[0][g];
          }
        }
        __syncthreads();
        double sg = 0.0;
        for (int e = tid; e < N; e += nt) {
          latent_local[e] = 1.0 / rng.gamma(P0 + 2, (uint32_t)e, 1.0, 1.0 / (1 + 1 / (local[e] * local[e])));
          const double lh = gsm[0][gi[e]] * local[e];
          sg += coef[e] * coef[e] / (2 * lh * lh);
        }
        double glob = 1.0;
        if (D.ggl) {
          sg = B.sum(sg);
          __shared__ double gnew;
          if (tid == 0) {
            const double lg = 1.0 / rng.gamma(P0 + 3, 0, 1.0, 1.0 / (1 + 1 / (glob_old * glob_old)));
            hs[HS_LATENT_GLOBAL] = lg;
            gnew = hs_global(rng, P0 + 4, 0, N, lg, sg);  // floor((N + 1) / 2)  QUIRK draw.h:672
            hs[HS_GLOBAL] = gnew;
          }
          __syncthreads();
          glob = gnew;
        }
        for (int e = tid; e < N; e += nt) {  // horseshoe_local_sparsity draw.h:649-656 with prior_var = global^2
          const double cv = gsm[0][gi[e]];
          const double scl = 1.0 / (1 / latent_local[e] + coef[e] * coef[e] / (2 * glob * glob * cv * cv));
          const double lv = sqrt(1.0 / rng.gamma(P0 + 5, (uint32_t)e, 1.0, scl));
          local[e] = lv;
          const double sdv = glob * cv * lv;
          const double pr = 1.0 / (sdv * sdv);
          put_prec(D, prec_u, e, pr);
          shrink[e] = 1.0 / (1 + pr);
        }
        break;
      }
      case 5: {  // Normal-Gamma, tri.h:793-812
        double* local = hN0;
        double* group = hG0;
        double* lshape = hG1;
        const double glob_old = D.ggl ? hs[HS_GLOBAL] : 1.0;
        for (int g = 0; g < G; ++g) {
          double slog = 0.0, ssq = 0.0;
          for (int e = tid; e < N; e += nt)
            if (gi[e] == g) { slog += log(local[e]); ssq += local[e] * local[e]; }
          slog = B.sum(slog);
          ssq = B.sum(ssq);
          if (tid == 0) {
            const int cnt = P.grp_cnt[g];
            const double gs = ng_shape_jump(rng, P0 + 0, P0 + 1, (uint32_t)g, lshape[g], cnt, slog, ssq, glob_old * group[g], C.ng_shape_sd);
            lshape[g] = gs;
            // ng_mn_sparsity draw.h:1076-1095; the scale passed is global_scale (tri.h:762)  QUIRK
            const double t = cut_param(sqrt(1.0 / rng.gamma(P0 + 2, (uint32_t)g, C.ng_group_shape + cnt * gs,
                                                            1.0 / (gs * ssq / (glob_old * glob_old) + C.ng_global_scale))));
            group[g] = t;
            gsm[0][g] = t;
            gsm[1][g] = gs;
          }
        }
        __syncthreads();
        double glob = 1.0;
        if (D.ggl) {  // ng_global_sparsity (vector overload) draw.h:1056-1070
          double sgam = 0.0, ssq = 0.0;
          for (int e = tid; e < N; e += nt) {
            const double gs = gsm[1][gi[e]], v = local[e] / gsm[0][gi[e]];
            sgam += gs;
            ssq += gs * v * v;
          }
          sgam = B.sum(sgam);
          ssq = B.sum(ssq);
          __shared__ double gnew;
          if (tid == 0) {
            gnew = cut_param(sqrt(1.0 / rng.gamma(P0 + 3, 0, C.ng_global_shape + sgam, 1.0 / (ssq + C.ng_global_scale))));
            hs[HS_GLOBAL] = gnew;
          }
          __syncthreads();
          glob = gnew;
        }
        for (int e = tid; e < N; e += nt) {  // ng_local_sparsity draw.h:1020-1031
          const double gs = gsm[1][gi[e]], gl = glob * gsm[0][gi[e]];
          const double t = cut_param(sqrt(rng.gig(P0 + 4, (uint32_t)e, gs - .5, 2 * gs / (gl * gl), coef[e] * coef[e])));
          local[e] = t;
          put_prec(D, prec_u, e, 1.0 / (t * t));
        }
        break;
      }
      case 6: {  // Dirichlet-Laplace, tri.h:885-901
        double* local = hN0;
        double* latent = hN1;
        double* group = hG0;
        const double glob_old = D.ggl ? hs[HS_GLOBAL] : 1.0;
        for (int g = 0; g < G; ++g) {  // dl_mn_sparsity (first overload) draw.h:808-834
          double sm = 0.0;
          for (int e = tid; e < N; e += nt)
            if (gi[e] == g) sm += fabs(coef[e]) / (glob_old * local[e]);
          sm = B.sum(sm);
          if (tid == 0) {
            const double t = cut_param(1.0 / rng.gamma(P0 + 0, (uint32_t)g, C.dl_shape + P.grp_cnt[g], 1.0 / (C.dl_scale + sm)));
            group[g] = t;
            gsm[0][g] = t;
          }
        }
        double slog = 0.0;
        for (int e = tid; e < N; e += nt) slog += log(local[e]);
        slog = B.sum(slog);
        const int grid = C.dl_grid;  // dl_dir_griddy draw.h:866-890
        for (int i = tid; i < grid; i += nt) {
          const double a = dl_grid_point(grid, N, i);
          gridbuf[i] = a * (N * (log(glob_old) - log(2.0)) + slog) - N * lgamma(a);
        }
        const double dir = dl_grid_point(grid, N, B.griddy_pick(rng, P0 + 1, grid));
        double sm = 0.0;
        for (int e = tid; e < N; e += nt) {  // dl_local_sparsity draw.h:778-785
          const double t = cut_param(rng.gig(P0 + 2, (uint32_t)e, dir - 1, 1.0, 2 * fabs(coef[e] / gsm[0][gi[e]])));
          local[e] = t;
          sm += t;
        }
        sm = B.sum(sm);
        double sg = 0.0;
        for (int e = tid; e < N; e += nt) {
          const double t = local[e] / sm;
          local[e] = t;
          sg += fabs(coef[e]) / (t * gsm[0][gi[e]]);
        }
        double glob = 1.0;
        if (D.ggl) {  // dl_global_sparsity draw.h:793-799
          sg = B.sum(sg);
          __shared__ double gnew;
          if (tid == 0) {
            gnew = cut_param(rng.gig(P0 + 3, 0, N * (dir - 1), 1.0, 2 * sg));
            hs[HS_GLOBAL] = gnew;
          }
          __syncthreads();
          glob = gnew;
        }
        for (int e = tid; e < N; e += nt) {  // dl_latent draw.h:758-770
          const double sc = glob * local[e] * gsm[0][gi[e]];
          const double t = cut_param(1.0 / rng.invgauss(P0 + 4, (uint32_t)e, sc / fabs(coef[e]), 1.0));
          latent[e] = t;
          put_prec(D, prec_u, e, 1.0 / (sc * sc * t));
        }
        if (tid == 0) hs[HS_DIR] = dir;
        break;
      }
      case 7: {  // GDP, tri.h:974-986
        double* local = hN0;
        double* grate = hG0;
        const double rate_old = hs[HS_GRATE];
        double sl = 0.0;
        sl = B.sum(log1p_sum(coef, N, tid, nt, 1.0 / rate_old));  // signed coefficient  QUIRK draw.h:1199
        int grid = C.gdp_grid_shape;
        for (int i = tid; i < grid; i += nt) {
          const double c = interior_grid(grid, i);
          gridbuf[i] = N * (log(1 - c) - log(c)) - sl / c;
        }
        double c = interior_grid(grid, B.griddy_pick(rng, P0 + 0, grid));
        const double shape = (1 - c) / c;
        grid = C.gdp_grid_rate;  // gdp_rate_griddy draw.h:1225-1235: one warp per grid point
        for (int g = tid >> 5; g < grid; g += nt >> 5) {
          const double cg = interior_grid(grid, g);
          const double s = wsum(log1p_sum(coef, N, tid & 31, 32, cg / (1 - cg)));
          if ((tid & 31) == 0) gridbuf[g] = N * (log(cg) - log(1 - cg)) - (shape + 1) * s;
        }
        c = interior_grid(grid, B.griddy_pick(rng, P0 + 1, grid));
        const double rate = (1 - c) / c;
        for (int g = 0; g < G; ++g) {  // gdp_exp_rate (group overload) draw.h:1173-1190
          double l1 = 0.0;
          for (int e = tid; e < N; e += nt)
            if (gi[e] == g) l1 += fabs(coef[e]);
          l1 = B.sum(l1);
          if (tid == 0) {
            gsm[0][g] = rng.gamma(P0 + 2, (uint32_t)g, shape + P.grp_cnt[g], 1.0 / (rate + l1));
            grate[g] = gsm[0][g];
          }
        }
        __syncthreads();
        for (int e = tid; e < N; e += nt) {  // gdp_local_sparsity draw.h:1147-1153
          const double r = gsm[0][gi[e]];
          const double t = 1.0 / rng.invgauss(P0 + 3, (uint32_t)e, fabs(r / coef[e]), r * r);
          local[e] = t;
          put_prec(D, prec_u, e, 1.0 / t);
        }
        if (tid == 0) { hs[HS_GSHAPE] = shape; hs[HS_GRATE] = rate; }
        break;
      }
    }
    return;
  }

  // ---------------- side 1: updateImpactPrec ----------------
  const double* a = P.contem + (size_t)u * q;
  double* cp = P.chol_prec + (size_t)u * q;
  double* hq0 = P.hq0 + (size_t)u * q;
  double* hq1 = P.hq1 + (size_t)u * q;
  double* hq2 = P.hq2 + (size_t)u * q;
  switch (D.prior) {
    case 1: break;
    case 4: {  // tri.h:525-531
      const double lam = hs[HS_CONTEM_LAMBDA];
      double chi = 0.0;
      for (int e = tid; e < q; e += nt) chi += a[e] * a[e] * cp[e] * lam;
      chi = B.sum(chi);
      __shared__ double newlam;
      if (tid == 0) {
        newlam = cut_param(rng.gig(P1 + 0, 0, C.hmn_shape - (double)(q / 2), 2 * C.hmn_rate, chi));
        hs[HS_CONTEM_LAMBDA] = newlam;
      }
      __syncthreads();
      for (int e = tid; e < q; e += nt) cp[e] = cp[e] * lam / newlam;
      break;
    }
    case 2: {  // tri.h:617-623
      double* dummy = hq0;
      double* slab = hq1;
      double* weight = hq2;
      const double spike = hs[HS_CONTEM_SPIKE];
      double ss = 0.0;
      for (int e = tid; e < q; e += nt) {
        const double s = sqrt(1.0 / rng.gamma(P1 + 0, (uint32_t)e, C.ssvs_chol_slab_shape + .5,
                                              1.0 / (C.ssvs_chol_slab_scl + a[e] * a[e] / (dummy[e] + (1 - dummy[e]) * spike))));
        slab[e] = s;
        const double r = a[e] / s;
        ss += r * r;
      }
      ss = B.sum(ss);
      const int grid = C.ssvs_chol_grid;
      for (int i = tid; i < grid; i += nt) {
        const double c = interior_grid(grid, i);
        gridbuf[i] = -ss / (2 * c * c) - q * log(c);
      }
      const double nspike = interior_grid(grid, B.griddy_pick(rng, P1 + 1, grid));
      double sm = 0.0;
      for (int e = tid; e < q; e += nt) {
        const double sn = slab[e], sdn = nspike * slab[e], wt = weight[e];
        double e1 = -a[e] * a[e] / (2 * sn * sn), e2 = -a[e] * a[e] / (2 * sdn * sdn);
        const double mx = fmax(e1, e2);
        e1 = wt * exp(e1 - mx) / sn;
        e2 = (1 - wt) * exp(e2 - mx) / sdn;
        const double dm = rng.bernoulli(P1 + 2, (uint32_t)e, e1 / (e1 + e2));
        dummy[e] = dm;
        sm += dm;
        const double sdv = (1 - dm) * sdn + dm * sn;  // build_ssvs_sd draw.h:112-116, tri.h:622
        cp[e] = 1.0 / (sdv * sdv);
      }
      sm = B.sum(sm);
      for (int e = tid; e < q; e += nt)  // ssvs_weight draw.h:290-297: one Beta draw per element
        weight[e] = rng.beta(P1 + 3, (uint32_t)e, C.ssvs_chol_s1 + sm, C.ssvs_chol_s2 + q - sm);
      if (tid == 0) hs[HS_CONTEM_SPIKE] = nspike;
      break;
    }
    case 3: {  // tri.h:712-720
      double* local = hq0;
      double* latent = hq1;
      const double glob_old = hs[HS_CONTEM_GLOBAL];
      double s = 0.0;
      for (int e = tid; e < q; e += nt) {
        const double lt = 1.0 / rng.gamma(P1 + 0, (uint32_t)e, 1.0, 1.0 / (1 + 1 / (local[e] * local[e])));
        latent[e] = lt;
        const double scl = 1.0 / (1 / lt + a[e] * a[e] / (2 * glob_old * glob_old));
        const double lv = sqrt(1.0 / rng.gamma(P1 + 2, (uint32_t)e, 1.0, scl));
        local[e] = lv;
        s += a[e] * a[e] / (2 * lt * lt);  // QUIRK tri.h:717: the latent vector is passed as the local scales
        const double sdv = glob_old * lv;  // QUIRK tri.h:715, 719: previous sweep's global
        cp[e] = 1.0 / (sdv * sdv);
      }
      s = B.sum(s);
      if (tid == 0) {
        const double lg = 1.0 / rng.gamma(P1 + 1, 0, 1.0, 1.0 / (1 + 1 / (glob_old * glob_old)));
        hs[HS_LATENT_CONTEM_GLOBAL] = lg;
        hs[HS_CONTEM_GLOBAL] = hs_global(rng, P1 + 3, 0, q, lg, s);
      }
      break;
    }
    case 5: {  // tri.h:813-818
      double* fac = hq0;
      const double glob_old = hs[HS_CONTEM_GLOBAL];
      double slog = 0.0, ssq = 0.0;
      for (int e = tid; e < q; e += nt) { slog += log(fac[e]); ssq += fac[e] * fac[e]; }
      slog = B.sum(slog);
      ssq = B.sum(ssq);
      __shared__ double sh_shape, sh_glob;
      if (tid == 0) {
        sh_shape = ng_shape_jump(rng, P1 + 0, P1 + 1, 0, hs[HS_CONTEM_SHAPE], q, slog, ssq, glob_old, C.ng_shape_sd);
        sh_glob = cut_param(sqrt(1.0 / rng.gamma(P1 + 2, 0, C.ng_cg_shape + q * sh_shape, 1.0 / (sh_shape * ssq + C.ng_cg_scale))));
        hs[HS_CONTEM_SHAPE] = sh_shape;
        hs[HS_CONTEM_GLOBAL] = sh_glob;
      }
      __syncthreads();
      for (int e = tid; e < q; e += nt) {
        const double t = cut_param(sqrt(rng.gig(P1 + 3, (uint32_t)e, sh_shape - .5, 2 * sh_shape / (sh_glob * sh_glob), a[e] * a[e])));
        fac[e] = t;
        cp[e] = 1.0 / (t * t);
      }
      break;
    }
    case 6: {  // tri.h:902-908
      double* local = hq0;
      double* latent = hq1;
      const double glob_old = hs[HS_CONTEM_GLOBAL];
      double slog = 0.0;
      for (int e = tid; e < q; e += nt) slog += log(local[e]);
      slog = B.sum(slog);
      const int grid = C.dl_grid;
      for (int i = tid; i < grid; i += nt) {
        const double aa = dl_grid_point(grid, q, i);
        gridbuf[i] = aa * (q * (log(glob_old) - log(2.0)) + slog) - q * lgamma(aa);
      }
      const double dir = dl_grid_point(grid, q, B.griddy_pick(rng, P1 + 0, grid));
      double sm = 0.0;
      for (int e = tid; e < q; e += nt) {
        const double t = cut_param(rng.gig(P1 + 1, (uint32_t)e, dir - 1, 1.0, 2 * fabs(a[e])));
        local[e] = t;
        sm += t;
      }
      sm = B.sum(sm);
      double sg = 0.0;
      for (int e = tid; e < q; e += nt) {
        const double t = local[e] / sm;
        local[e] = t;
        sg += fabs(a[e]) / t;
      }
      sg = B.sum(sg);
      __shared__ double gnew;
      if (tid == 0) {
        gnew = cut_param(rng.gig(P1 + 2, 0, q * (dir - 1), 1.0, 2 * sg));
        hs[HS_CONTEM_GLOBAL] = gnew;
        hs[HS_CONTEM_DIR] = dir;
      }
      __syncthreads();
      for (int e = tid; e < q; e += nt) {  // QUIRK tri.h:906: the latent uses the local scale only
        const double t = cut_param(1.0 / rng.invgauss(P1 + 3, (uint32_t)e, local[e] / fabs(a[e]), 1.0));
        latent[e] = t;
        const double sc = gnew * local[e];
        cp[e] = 1.0 / (sc * sc * t);
      }
      break;
    }
    case 7: {  // tri.h:987-993
      double* fac = hq0;
      double* crate = hq1;
      const double rate_old = hs[HS_CGRATE];
      double sl = 0.0;
      sl = B.sum(log1p_sum(a, q, tid, nt, 1.0 / rate_old));
      int grid = C.gdp_grid_shape;
      for (int i = tid; i < grid; i += nt) {
        const double c = interior_grid(grid, i);
        gridbuf[i] = q * (log(1 - c) - log(c)) - sl / c;
      }
      double c = interior_grid(grid, B.griddy_pick(rng, P1 + 0, grid));
      const double shape = (1 - c) / c;
      grid = C.gdp_grid_rate;
      for (int g = tid >> 5; g < grid; g += nt >> 5) {
        const double cg = interior_grid(grid, g);
        const double s = wsum(log1p_sum(a, q, tid & 31, 32, cg / (1 - cg)));
        if ((tid & 31) == 0) gridbuf[g] = q * (log(cg) - log(1 - cg)) - (shape + 1) * s;
      }
      c = interior_grid(grid, B.griddy_pick(rng, P1 + 1, grid));
      const double rate = (1 - c) / c;
      for (int e = tid; e < q; e += nt) {
        const double r = rng.gamma(P1 + 2, (uint32_t)e, shape + 1, 1.0 / (rate + fabs(a[e])));  // gdp_exp_rate draw.h:1160-1165
        crate[e] = r;
        const double t = 1.0 / rng.invgauss(P1 + 3, (uint32_t)e, fabs(r / a[e]), r * r);
        fac[e] = t;
        cp[e] = 1.0 / t;
      }
      if (tid == 0) { hs[HS_CGSHAPE] = shape; hs[HS_CGRATE] = rate; }
      break;
    }
  }
}

cudaError_t launch_prior(const Dims& D, const DevPtrs& P, const PriorConst& C, int side, cudaStream_t st) {
  if (D.prior == 1) return cudaSuccess;
  // one CTA per unit; large coefficient vectors (k = 100, d = 301: N = 30 000) get 16 warps
  const int threads = (side == 0 ? D.N : D.q) >= 8192 ? 512 : 256;
  // batches of at most two CTAs per SM run as one wave whatever the register budget: they get the 128-register instance
  // (no spills, shortest single-CTA latency - the strong-scaling case of 128 chains per GPU); larger batches the 64-register one
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const bool one_wave = D.U <= 2 * sms;
  if (threads == 512) {
    if (D.U <= sms) prior_kernel<512, 1><<<D.U, 512, 0, st>>>(D, P, C, side);
    else prior_kernel<512, 2><<<D.U, 512, 0, st>>>(D, P, C, side);
  } else if (one_wave) prior_kernel<256, 2><<<D.U, 256, 0, st>>>(D, P, C, side);
  else prior_kernel<256, 4><<<D.U, 256, 0, st>>>(D, P, C, side);
  return cudaGetLastError();
}

}  // namespace bvhar
