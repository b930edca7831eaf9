// engine.cuh — device-side data model of the Gibbs engine (shared by all kernels).
//
// Units: u = window * C + chain.  All chains of a window share the design; windows are row ranges
// of one design matrix built on the whole series (forecaster.h:851-875, 920-956).
//
// HBM layout (DESIGN.md "Data layout"):
//   * time x series matrices of every unit of a window (log-volatility H, latent innovations Z,
//     inverse variances Dinv, weighted responses YLD) are stored time-major ACROSS the window's
//     units: [t][m], m = chain * k + series, pitch ldM.  This is (a) the K-major operand layout of
//     the DMMA contractions and (b) perfectly coalesced for one-thread-per-series recurrences.
//   * per-unit vectors (coefficients, hyper-parameters, S/C sufficient statistics) are unit-major.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bvhar {

struct Dims {
  int W, C, U;          // windows, chains per window, units
  int k, d, N, nrow, q, G;
  int Pp, ldS;          // packed lower size d(d+1)/2 and its pitch
  int ldX, ldXX, ldT;   // pitches of X [T][ldX], XX [T][ldXX], XT [d][ldT]
  int M, ldM;           // C*k and its pitch
  int T;                // rows of the full design
  int mean, sv, ggl, prior;
  int nmax;             // largest window
  int ubase;            // global index of unit 0 (Philox key word 1 = ubase + u)
};

struct PriorConst {  // scalars of bvhar_prior_params
  double hmn_shape, hmn_rate;
  int hmn_grid;
  double ssvs_chol_s1, ssvs_chol_s2, ssvs_coef_slab_shape, ssvs_coef_slab_scl, ssvs_chol_slab_shape, ssvs_chol_slab_scl;
  int ssvs_coef_grid, ssvs_chol_grid;
  double ng_shape_sd, ng_group_shape, ng_group_scale, ng_global_shape, ng_global_scale, ng_cg_shape, ng_cg_scale;
  int dl_grid;
  double dl_shape, dl_scale;
  int gdp_grid_shape, gdp_grid_rate;
};

struct DevPtrs {
  // ---- shared, read-only ----
  const double* X;     // [T][ldX]
  const double* XT;    // [d][ldT]
  const double* Y;     // [T][k]
  const double* XX;    // [T][ldXX]  x_ta * x_tb, packed pairs a >= b   (SV only)
  const double* xnorm; // [W][d]     column squared norms of the window's X (draw.h:425)
  const double* XtX;   // [W][ldS]   packed X'X          (LDLT only)
  const double* XtY;   // [W][d][k]  X'Y row-major       (LDLT only)
  const int* win_row0; // [W]
  const int* win_n;    // [W]
  const long long* win_off;  // [W] element offset of the window's [t][m] block
  const int* grp_idx;  // [N] index of the coefficient's group in grp_id
  const int* grp_cnt;  // [G] group sizes
  const int* own_flag; // [N] 1 if the group is an own-lag group (alpha_penalty = 0, tri.h:254-262)
  const double* prior_mean;  // [k*d] internal (j*d + a) layout, constants included
  const double* sig_shp;     // [k]
  const double* sig_scl;     // [k]
  const double* sv_mean;     // [k]
  const double* sv_prec;     // [k][k]
  const double* ssvs_s1;     // [G]
  const double* ssvs_s2;     // [G]
  const uint32_t* seeds;     // [U]
  // ---- [t][m] state / scratch ----
  double* H;     // lvol_draw
  double* Z;     // latent_innov
  double* Dinv;  // 1 / sqrt_sv^2 ; reused as the tridiagonal pivots in updateState
  double* YLD;   // (Y L')_ti / sqrt_sv_ti^2 ; reused as the tridiagonal right-hand side
  double* Zn;    // standard normals of the log-volatility draw, pre-drawn per sweep (sv_normals_kernel)
  // ---- per unit ----
  double* S;      // [U*k][ldS]  S_i = X' D_i X packed
  double* Cm;     // [U*k][ldX]  C_i = X' D_i (Y L_i')
  double* Acoef;  // [W][d][ldM] coefficient matrix in contraction layout
  double* alpha;  // [U][N]      coef_vec.head(num_alpha)
  double* cvec;   // [U][k]      constants
  double* prec;   // [U][k*d]    prior_alpha_prec, internal (j*d + a) layout
  double* contem; // [U][q]
  double* L;      // [U][k][k]   chol_lower, row-major
  double* chol_prec;     // [U][q]
  double* sparse_coef;   // [U][d*k] column-major d x k
  double* sparse_contem; // [U][q]
  double* lvol_init;     // [U][k]
  double* lvol_sig;      // [U][k]
  double* diag;          // [U][k]
  double* znorm;         // [U][k] squared norms of latent columns (draw_savs)
  // ---- prior state (unit-major) ----
  double *hN0, *hN1, *hN2, *hN3;  // [U][N]
  double *hG0, *hG1, *hG2;        // [U][G]
  double *hq0, *hq1, *hq2, *hq3;  // [U][q]
  double* hs;                     // [U][16] scalars
  // ---- control ----
  uint32_t* iter;  // Philox iteration counter (device resident so that a captured graph can be replayed)
  int* rec_row;    // next record row
  int* flags;      // [U] numerical flags
};

// scalar slots in hs[u][.]
enum HsSlot {
  HS_GLOBAL = 0, HS_LATENT_GLOBAL = 1, HS_CONTEM_GLOBAL = 2, HS_LATENT_CONTEM_GLOBAL = 3,
  HS_SPIKE = 4, HS_CONTEM_SPIKE = 5, HS_OWN_LAMBDA = 6, HS_CROSS_LAMBDA = 7, HS_CONTEM_LAMBDA = 8,
  HS_CONTEM_SHAPE = 9, HS_DIR = 10, HS_CONTEM_DIR = 11, HS_GSHAPE = 12, HS_GRATE = 13, HS_CGSHAPE = 14, HS_CGRATE = 15,
  HS_SLOTS = 16
};

struct RecPtrs {  // record buffers [U][cap][cols]; nullptr when not recorded
  double *alpha, *c, *a, *d, *h, *h_last, *h0, *sigh, *alpha_sparse, *c_sparse, *a_sparse;
  double *gamma, *lambda, *eta, *tau, *kappa;
  int cap;
};

__device__ __forceinline__ int tri_idx(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b

}  // namespace bvhar
