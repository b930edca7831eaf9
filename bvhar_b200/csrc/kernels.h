// kernels.h — host-callable launch wrappers of the sweep kernels (definitions in kernels_*.cu).
#pragma once
#include <cuda_runtime.h>

#include "engine.cuh"

namespace bvhar {

cudaError_t launch_sv_prep(const Dims& D, const DevPtrs& P, cudaStream_t st);
cudaError_t launch_sv_finish(const Dims& D, const DevPtrs& P, cudaStream_t st);  // sigma_h^2 and h_0 (kernels_sv_finish.cu)
cudaError_t launch_sv_wmix(const Dims& D, const DevPtrs& P, double* Wt, cudaStream_t st);  // weights of the direct P_j build (large-design SV path)
bool coef_plan(const Dims& D, size_t smem_limit, int* CH_out, bool* vglobal_out, bool* pglobal_out);
cudaError_t launch_coef(const Dims& D, const DevPtrs& P, double* vscratch, double* pscratch, size_t smem_limit, cudaStream_t st);
cudaError_t launch_draw_coef_batched(long long batch, int d, const double* Pp, const double* b, const double* z, double* out,
                                     cudaStream_t st);
bool coef_v2_fits(const Dims& D, size_t smem_limit);
bool coef_pipe_fits(const Dims& D, size_t smem_limit);
cudaError_t launch_coef_seq_nc1(const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_seq_nc2(const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_seq_nc4(const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_v2(int part, const Dims& D, const DevPtrs& P, double* Pfac, double* Qn, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_chol_nc1(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_chol_nc2(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st);
cudaError_t launch_coef_chol_nc4(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st);
inline cudaError_t launch_coef_chol(int nc, const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st) {
  return nc <= 1 ? launch_coef_chol_nc1(D, P, Pfac, smem_limit, st)
       : nc <= 2 ? launch_coef_chol_nc2(D, P, Pfac, smem_limit, st)
                 : launch_coef_chol_nc4(D, P, Pfac, smem_limit, st);
}
bool coef_big_fits(const Dims& D, size_t smem_limit);
// Rres (part 2, LDLT): R = X'Y - X'X A of the sweep's initial coefficients ([W][d][ldM], from the engine's residual contraction) or nullptr
cudaError_t launch_coef_big(int part, const Dims& D, const DevPtrs& P, double* Pfac, double* hscr, size_t smem_limit, cudaStream_t st,
                            const double* Rres = nullptr);
cudaError_t launch_chol_big(const Dims& D, const DevPtrs& P, double* Pfac, size_t smem_limit, cudaStream_t st);
cudaError_t launch_seq_big_nc8(const Dims& D, const DevPtrs& P, const double* Pfac, double* hscr, size_t smem, cudaStream_t st, const double* Rres = nullptr);
cudaError_t launch_seq_big_nc16(const Dims& D, const DevPtrs& P, const double* Pfac, double* hscr, size_t smem, cudaStream_t st, const double* Rres = nullptr);
size_t seq_time_smem(const Dims& D);
size_t seq_time_scratch_doubles(const Dims& D);  // per-fit scratch of the time-domain sequential kernel (SV, large d)
cudaError_t launch_seq_time(const Dims& D, const DevPtrs& P, const double* Pfac, double* escr, size_t smem_limit, cudaStream_t st);
size_t chol_cta_smem(const Dims& D);
size_t seq_big_smem(const Dims& D);
int impact_gsize(const Dims& D);
size_t impact_smem(const Dims& D);
bool impact_supported(const Dims& D, size_t smem_limit);
cudaError_t launch_impact(const Dims& D, const DevPtrs& P, double* gscratch, cudaStream_t st);
int impact_launches(const Dims& D);  // kernels one launch_impact call starts
cudaError_t launch_sv_state(int part, const Dims& D, const DevPtrs& P, cudaStream_t st);  // 0 mix, 1 solve, 2 finish, 3 normals of the solve
cudaError_t launch_impact_rows_nc1(const Dims& D, const DevPtrs& P, const double* gscratch, int gsize, int kz, int jlo, int jhi, cudaStream_t st);
cudaError_t launch_impact_rows_nc2(const Dims& D, const DevPtrs& P, const double* gscratch, int gsize, int kz, int jlo, int jhi, cudaStream_t st);
cudaError_t launch_impact_rows_nc4(const Dims& D, const DevPtrs& P, const double* gscratch, int gsize, int kz, int jlo, int jhi, cudaStream_t st);
bool ldlt_own_gram(const Dims& D);  // small k: impact_ldlt_kernel forms Z'Z itself, no gram_Z contraction
cudaError_t launch_ldlt_gram_small(const Dims& D, const DevPtrs& P, double* gscratch, cudaStream_t st);
cudaError_t launch_ldlt_state(const Dims& D, const DevPtrs& P, const double* gscratch, cudaStream_t st);  // reads the Z'Z of gram_Z
cudaError_t launch_record(const Dims& D, const DevPtrs& P, const RecPtrs& R, cudaStream_t st);
cudaError_t launch_bump(const DevPtrs& P, int bump_iter, int bump_rec, cudaStream_t st);
// prior hyper-parameter updaters (kernels_prior.cu): side 0 = updateCoefPrec, 1 = updateImpactPrec
cudaError_t launch_prior(const Dims& D, const DevPtrs& P, const PriorConst& C, int side, cudaStream_t st);
// setup helpers (kernels_setup.cu)
cudaError_t launch_scatter_lvol(const double* src, double* dst, int n, int M, int ldM, int nsrc, cudaStream_t st);
cudaError_t launch_record_mean(const double* rec, int U, int cap, int cols, int rows, int thin, double* out, cudaStream_t st);
// posterior summaries of one record (kernels_summary.cu): out[col][8] = mean, sd, q(level / 2), median, q(1 - level / 2), split-R-hat, ESS, pip
cudaError_t launch_record_summary(const double* rec, int U, int cap, int cols, int S, int thin, double level, double* out, cudaStream_t st);
cudaError_t launch_build_xx(const Dims& D, const DevPtrs& P, double* XX, cudaStream_t st);
// stability filter (kernels_stable.cu)
constexpr int STAB_MAX_SQUARINGS = 40;
cudaError_t launch_stab_build(const double* coef, long long ld, long long n_mat, int k, int nrow, int p, const double* har, int ldh,
                              double inv_thr, double* R, double* Rt, int ldm, double* logacc, int* state, cudaStream_t st);
cudaError_t launch_stab_norm(double* R, double* Rt, long long n_mat, int m, int ldm, double* logacc, int* state, int* kg, int* undecided,
                             cudaStream_t st);
// spillover densities (kernels_spillover.cu)
struct SpilloverArgs {
  int k, nrow, lag, step, sv_kind, n_time, ldh;
  long long n_draws, ld_coef, ld_contem, ld_sv;
  long long per_unit, us_coef, us_contem, us_sv;  // draw i -> unit i / per_unit (stride us_*), row i % per_unit (stride ld_*)
  const double *coef, *contem, *sv, *har;  // har: (lag k) rows x ldh: har[r * ldh + q] = har_trans(q, r)
  double* scratch;
  size_t slab;
  double *connect, *to, *from, *tot, *net_pairwise;
};
size_t spillover_slab_doubles(int k, int lag, int step);
cudaError_t launch_spillover(const SpilloverArgs& A, int n_cta, cudaStream_t st);
// forecast (kernels_forecast.cu)
struct RecView {  // element (unit u, draw i, column c) = p[u * us + i * sd + c * sc]
  const double* p;
  long long us, sd, sc;
};
struct ForecastArgs {
  int n_units, k, nrow, mean, sv_model, step, p, is_vhar, sv, get_lpl, S, har_r, har_c;
  unsigned unit_base;  // Philox key word 1 of unit u = unit_base + u
  const double* har;
  RecView coef, cst, contem, diag, sigh;  // alpha | constants | a | d or h(last) | sigh
  const double *last_obs, *valid;
  const uint32_t* seed;
  double *out, *lpl;
  double* lpl_draw;   // [n_units][step][S] per-draw log predictive likelihoods (scratch owned by launch_forecast)
  const double* act;  // [n_units][ncoef] activity graph of the Select forecasters (0 or 1.1), nullptr for the plain ones
  int ncoef;          // nrow * k + k * mean
};
cudaError_t launch_forecast(const ForecastArgs& A, cudaStream_t st);
// RegRecords::computeActivity (config.h:747-758) of every unit: act[u][i] for the ncoef = N + kc record columns
// (alpha, then the constants); S draws per unit; n_lo / n_hi = the 1-based order statistics of the two tail quantiles
size_t activity_scratch_doubles(int n_units, int ncoef, int S);
cudaError_t launch_activity(const RecView& coef, const RecView& cst, int n_units, int N, int kc, int S, int n_lo, int n_hi, double* act,
                            double* scratch, cudaStream_t st);

// conjugate Minnesota samplers (kernels_mniw.cu): McmcMniw / MhMinnesota, bayes/mniw/minnesota.h:241-278, 413-510
struct MniwArgs {
  int k, d, chains, rows, num_burn, thin;
  double shape;          // posterior inverse-Wishart shape
  const double* mean;    // [k][d]: d x k posterior mean, column-major
  const double* Rfull;   // [d][d] row-major upper factor R of the posterior precision (prec = R'R)
  const double* Ls;      // [k][k] row-major lower Cholesky factor of the inverse-Wishart scale
  const uint32_t* seeds; // [chains]
  uint32_t ubase;
  double* wscratch;      // nullptr: the d x k work matrix lives in shared memory; else [grid][d * k]
  double* alpha;         // [chains][d * k][rows] (column-major rows x (d k) per chain)
  double* sigma;         // [chains][k * k][rows]
  int* flags;            // bit 0: a sampled Sigma was not positive definite
};
struct MhArgs {
  int k, chains, rows, num_iter, num_burn, thin;
  double gamma_shp, gamma_rate, invgam_shp, invgam_scl;  // as MhMinnesota stores them (minnesota.h:426-427)
  const double* init_par;   // [chains][1 + k] prevprior = (lambda, psi)
  const double* chol_var;   // [chains][(1 + k)^2] row-major upper Cholesky factor of scale_variance * hessian^-1
  const uint32_t* seeds;
  uint32_t ubase;
  unsigned char* accept_all;  // [chains][num_iter] scratch
  double* lambda;             // [chains][rows]
  double* psi;                // [chains][k][rows]
  unsigned char* accept;      // [chains][rows]
  int* flags;                 // bit 1: a proposed psi was negative (the reference STOPs)
};
size_t mniw_smem(int k, int d, bool w_in_smem);
cudaError_t launch_mniw(const MniwArgs& A, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_mh(const MhArgs& M, cudaStream_t st);

}  // namespace bvhar
