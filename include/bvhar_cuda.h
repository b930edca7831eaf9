/*
 * bvhar_cuda.h — C-ABI of the B200-native Gibbs engine that replaces the hot path of
 * ygeunkim/bvhar (the corrected-triangular sampler behind var_bayes()/vhar_bayes() and the
 * forecast_roll()/forecast_expand() refit loops).
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, never throws, and returns
 * 0 on success or a bvhar_status error class; the message is available from
 * bvhar_cuda_last_error() (thread-local).  All matrices are column-major fp64 (Eigen / R /
 * Fortran-ordered NumPy), all integer vectors are int32, all pointers are caller-owned HOST
 * memory that only has to stay valid for the duration of the call it is passed to.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   bvhar_cuda_fit_*          <- bvhar::McmcRun<BaseMcmc,isGroup> ctor + returnRecords()
 *                                inst/include/bvhar/src/bayes/triangular/triangular.h:1192-1295
 *                                (bound by src/triangular.cpp:26-70 `estimate_sur` and
 *                                 python/src/bvhar/_src/_cta.cpp:4-38 McmcLdlt/SvMcmc/...)
 *   bvhar_cuda_outforecast_*  <- bvhar::McmcVarforecastRun / McmcVharforecastRun
 *                                <McmcRollforecastRun|McmcExpandforecastRun, RegForecaster|SvForecaster>
 *                                inst/include/bvhar/src/bayes/triangular/forecaster.h:575-1121
 *                                (bound by src/triangular.cpp:159-630 roll_* / expand_* and
 *                                 _cta.cpp:40-208 LdltVarRoll ... SvGrpVharExpand)
 *   bvhar_cuda_forecast_*     <- bvhar::McmcForecastRun<BaseForecaster> forecaster.h:494-568
 *   record names              <- RegRecords/LdltRecords/SvRecords/SparseRecords + prior records,
 *                                inst/include/bvhar/src/bayes/triangular/config.h:581-1251
 */
#ifndef BVHAR_CUDA_H
#define BVHAR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVHAR_CUDA_ABI_VERSION 2

typedef enum {
  BVHAR_OK = 0,
  BVHAR_ERR_BAD_ARG = 1,     /* inconsistent descriptor (maps to STOP()/value_error in the adapter) */
  BVHAR_ERR_CUDA = 2,        /* CUDA runtime failure, no device, extension not built for this GPU */
  BVHAR_ERR_NUMERICAL = 3,   /* non-finite state detected in some unit (flagged, not aborted) */
  BVHAR_ERR_INTERRUPTED = 4, /* progress callback asked to stop; partial records are available */
  BVHAR_ERR_UNSUPPORTED = 5  /* valid request that this build does not cover (size limits) */
} bvhar_status;

/* config.h:50-102 — `param_reg` presence of "initial_mean" selects SV (src/triangular.cpp:33). */
typedef enum { BVHAR_COV_LDLT = 0, BVHAR_COV_SV = 1 } bvhar_cov_type;

/* triangular.h:1043-1165 `prior_type` switch. */
typedef enum {
  BVHAR_PRIOR_MINNESOTA = 1,
  BVHAR_PRIOR_SSVS = 2,
  BVHAR_PRIOR_HORSESHOE = 3,
  BVHAR_PRIOR_HIERMINN = 4,
  BVHAR_PRIOR_NG = 5,
  BVHAR_PRIOR_DL = 6,
  BVHAR_PRIOR_GDP = 7
} bvhar_prior_type;

/* What to keep on the device per posterior draw.  The reference stores every draw of every
 * parameter (config.h:587-589, 924-928); h_record is n*k doubles per draw, so callers with many
 * chains can drop it (the forecaster only needs the last time point, config.h:998-1001). */
enum {
  BVHAR_REC_COEF = 1u << 0,   /* alpha/c/a + d | h0,sigh (always cheap) */
  BVHAR_REC_SPARSE = 1u << 1, /* alpha_sparse/c_sparse/a_sparse (SAVS draws) */
  BVHAR_REC_PRIOR = 1u << 2,  /* gamma | lambda,eta,tau,kappa ... per prior */
  BVHAR_REC_LVOL = 1u << 3,   /* full h_record (n*k per draw) */
  BVHAR_REC_LVOL_LAST = 1u << 4, /* only h at the last time point (k per draw): enough to forecast */
  BVHAR_REC_ALL = 0x0Fu
};

/* Shrinkage-prior hyper-parameters: the union of the `param_prior` keys parsed in
 * config.h:124-143 (Minnesota), 180 (hierarchical), 243-248 (SSVS), 298-301 (NG), 327 (DL),
 * 352 (GDP).  Horseshoe has none (config.h:257-269). */
typedef struct {
  /* Minnesota / hierarchical Minnesota: moments already reduced by the host adapter from the
   * dummy observations (config.h:144-151, design.h:60-98): length num_alpha each, alpha order
   * (equation-major, constants excluded).  prec = diag(kron(diag(1/sigma), Xd'Xd)). */
  const double* minn_prior_mean;
  const double* minn_prior_prec;
  double hmn_shape, hmn_rate; /* config.h:180 */
  int32_t hmn_grid_size;
  /* SSVS */
  const double* ssvs_coef_s1; /* n_grp */
  const double* ssvs_coef_s2; /* n_grp */
  double ssvs_chol_s1, ssvs_chol_s2;
  double ssvs_coef_slab_shape, ssvs_coef_slab_scl, ssvs_chol_slab_shape, ssvs_chol_slab_scl;
  int32_t ssvs_coef_grid, ssvs_chol_grid;
  /* NG */
  double ng_shape_sd, ng_group_shape, ng_group_scale, ng_global_shape, ng_global_scale;
  double ng_contem_global_shape, ng_contem_global_scale;
  /* DL */
  int32_t dl_grid_size;
  double dl_shape, dl_scale;
  /* GDP */
  int32_t gdp_grid_shape, gdp_grid_rate;
} bvhar_prior_params;

/* One chain's initial state: the union of the `param_init[chain]` keys parsed in
 * config.h:371-373, 383-385, 410-420, 434-440, 457-475, 491-503, 515-521, 534-542, 558-574.
 * Unused pointers are NULL. */
typedef struct {
  const double* init_coef;   /* dim_design x dim, column-major */
  const double* init_contem; /* q = dim(dim-1)/2 */
  const double* init_diag;   /* LDLT: dim */
  const double* lvol_init;   /* SV: dim */
  const double* lvol;        /* SV: n x dim column-major (n = rows of window 0); may be NULL when
                                `lvol_from_init` is set (config.h:416-420) */
  const double* lvol_sig;    /* SV: dim */
  double own_lambda, cross_lambda, contem_lambda; /* hierarchical Minnesota */
  const double* coef_dummy;   /* SSVS init_coef_dummy: num_alpha */
  const double* coef_mixture; /* SSVS: n_grp */
  const double* chol_mixture; /* SSVS: q */
  const double* coef_slab;    /* SSVS: num_alpha */
  const double* contem_slab;  /* SSVS: q */
  double coef_spike_scl, chol_spike_scl;
  const double* local_sparsity;        /* HS/NG/DL/GDP: num_alpha */
  double global_sparsity;              /* HS/NG/DL */
  const double* group_sparsity;        /* HS/NG: n_grp */
  const double* contem_local_sparsity; /* q */
  double contem_global_sparsity;       /* length-1 vector in the reference */
  const double* local_shape; /* NG: n_grp */
  double contem_shape;       /* NG */
  const double* group_rate;  /* GDP: n_grp */
  const double* contem_rate; /* GDP: q */
  double gamma_shape, gamma_rate, contem_gamma_shape, contem_gamma_rate; /* GDP */
} bvhar_chain_init;

/* A batch of independent sampling units = windows x chains, all sharing one design/response
 * pair; window w uses rows [win_row0[w], win_row0[w] + win_nrow[w]).  A plain
 * McmcRun fit is n_windows = 1, win_row0 = {0}, win_nrow = {n}.  Rolling / expanding windows are
 * row ranges of the design built once on the whole series (forecaster.h:851-859, 920-928). */
typedef struct {
  int32_t abi_version; /* BVHAR_CUDA_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal this engine runs on */
  int32_t n_windows;
  int32_t n_chains; /* per window */
  int32_t dim;        /* k */
  int32_t dim_design; /* d = k*p (+1) */
  int32_t include_mean;
  int32_t cov_type;   /* bvhar_cov_type */
  int32_t prior_type; /* bvhar_prior_type */
  int32_t is_group;   /* template argument isGroup (`ggl`) */
  int32_t num_iter;   /* total sweeps, burn-in included (triangular.h:1195) */
  int32_t num_burn;
  int32_t thin;
  uint32_t record_mask;
  int32_t lvol_from_init; /* SvInits(init, num_design): replicate lvol_init over time (config.h:416-420) */
  int32_t unit_id_base; /* global index of this engine's first unit: unit u draws from the Philox stream keyed
                           (seed_chain[u], unit_id_base + u), so a batch split over several engines / GPUs
                           reproduces the draws of the unsplit batch bit for bit */
  int64_t n_rows_total; /* leading dimension of x and y */
  const double* x;      /* n_rows_total x dim_design */
  const double* y;      /* n_rows_total x dim */
  const int32_t* win_row0;
  const int32_t* win_nrow;
  /* Minnesota-type group structure (config.h:78-79, draw.h:75-82) */
  const int32_t* grp_mat; /* (num_alpha/dim) x dim column-major -> grp_vec = reshaped() */
  const int32_t* grp_id;
  int32_t n_grp;
  int32_t n_own;
  const int32_t* own_id;
  const int32_t* cross_id; /* accepted and ignored, as in the reference (set_grp_id by value, draw.h:75) */
  int32_t n_cross;
  int32_t reserved1;
  /* covariance prior (config.h:71-72, 100-101) */
  const double* sig_shape; /* dim */
  const double* sig_scale; /* dim */
  const double* sv_init_mean; /* SV: dim */
  const double* sv_init_prec; /* SV: dim x dim */
  /* intercept prior (config.h:73-74) */
  const double* mean_non; /* dim */
  double sd_non;
  bvhar_prior_params prior;
  const bvhar_chain_init* init; /* n_chains entries, reused by every window (forecaster.h:867-872) */
  const uint32_t* seed_chain;   /* n_windows x n_chains, row-major by window (forecaster.h:871) */
} bvhar_fit_desc;

typedef struct bvhar_fit bvhar_fit;

/* Return non-zero from the callback to interrupt (bvharinterrupt + 5 % logging,
 * triangular.h:1248-1275). `done`/`total` count sweeps of the whole batch. */
typedef int (*bvhar_progress_fn)(void* ctx, int done, int total);

const char* bvhar_cuda_last_error(void);
int bvhar_cuda_abi_version(void);
int bvhar_cuda_device_count(int* count);
/* sizeof() of the descriptor structs as compiled, so that FFI mirrors can verify their layout. */
int bvhar_cuda_struct_sizes(int64_t* fit_desc, int64_t* chain_init, int64_t* forecast_desc, int64_t* prior_params);
int bvhar_cuda_outforecast_params_size(int64_t* bytes);

/* Page-locked host buffers for record / forecast outputs (cudaHostAlloc): device-to-host copies into them run at
 * PCIe speed instead of the pageable staging path.  Optional: every entry point accepts ordinary memory too. */
int bvhar_cuda_host_alloc(int64_t bytes, void** out);
int bvhar_cuda_host_free(void* p);

int bvhar_cuda_fit_create(const bvhar_fit_desc* desc, bvhar_fit** out);
/* Run num_burn warm-up sweeps and num_iter - num_burn recorded sweeps for every unit
 * (McmcRun::fit / runGibbs, triangular.h:1217-1284). */
int bvhar_cuda_fit_run(bvhar_fit* fit, bvhar_progress_fn progress, void* ctx);
/* Finer-grained driver used by benches and tests: `n_sweeps` more sweeps, recorded or not. */
int bvhar_cuda_fit_step(bvhar_fit* fit, int n_sweeps, int record);
/* Block until every queued sweep has finished; reports per-unit numerical flags. */
int bvhar_cuda_fit_sync(bvhar_fit* fit);
/* Device time (ms, CUDA events on the engine's stream) of the last fit_step / fit_run call. */
int bvhar_cuda_fit_last_ms(bvhar_fit* fit, double* ms);
/* Per-kernel timing: while on, sweeps are launched eagerly with a CUDA-event pair around every kernel
 * (the sv_state entry covers its three kernels); bvhar_cuda_fit_profile writes one line per kernel,
 * "name launches total_ms". */
int bvhar_cuda_fit_set_profiling(bvhar_fit* fit, int on);
int bvhar_cuda_fit_profile(bvhar_fit* fit, char* buf, int64_t buflen);
/* Number of kernels launched by this engine since creation. */
int bvhar_cuda_fit_launch_count(bvhar_fit* fit, int64_t* launches);

/* Record access.  Names follow config.h: alpha_record, c_record, a_record, d_record, h_record,
 * h0_record, sigh_record, alpha_sparse_record, c_sparse_record, a_sparse_record, gamma_record,
 * lambda_record, eta_record, tau_record, kappa_record.  Thinning (draw.h:1297-1310) is applied
 * here; `rows` = kept draws. `out` is column-major rows x cols. */
int bvhar_cuda_fit_record_dims(bvhar_fit* fit, const char* name, int64_t* rows, int64_t* cols);
int bvhar_cuda_fit_record(bvhar_fit* fit, int window, int chain, const char* name, double* out,
                          int64_t rows, int64_t cols);
/* All units of one record in a single call: out[unit][kept draw][column], C order, unit = window *
 * n_chains + chain. */
int bvhar_cuda_fit_record_all(bvhar_fit* fit, const char* name, double* out, int64_t units, int64_t rows, int64_t cols);
/* Optional: bind `host` (units x rows x cols doubles, rows = final kept draws, same layout as _record_all; page-locked
 * memory makes the copies asynchronous) as the sink of one record BEFORE bvhar_cuda_fit_run.  The run then moves the
 * record's rows to the host on a copy stream while later sweeps are still sampling, and _record_all on the same pointer
 * only waits for the tail instead of copying: returnRecords() (triangular.h:1230-1233) without a serial device-to-host
 * pass at the end.  The buffer must stay valid until the fit is destroyed or the record has been fetched. */
int bvhar_cuda_fit_bind_record_sink(bvhar_fit* fit, const char* name, double* host, int64_t units, int64_t rows, int64_t cols);
/* Posterior mean of every column of one record over ALL units and kept draws (NaN entries of the SAVS records
 * skipped), reduced on the device: what the front ends compute from the returned draws (R/var-bayes.R:488-617
 * colMeans; _bayes.py coef_ / intercept_) without shipping the draws first. */
int bvhar_cuda_fit_record_mean(bvhar_fit* fit, const char* name, double* out, int64_t cols);
/* Names of the records this fit exports, '\n'-separated, in the reference's list order. */
int bvhar_cuda_fit_record_names(bvhar_fit* fit, char* buf, int64_t buflen);
/* Posterior summaries of one record over all chains and kept draws, reduced on the device (SURVEY.md section 8(f) row 1;
 * replaces the host passes of R/var-bayes.R:503-531 - colMeans, pip = colMeans(alpha_sparse != 0) - and of
 * posterior::summarise_draws on the returned draws).  out: `cols` x 8 row-major:
 * mean, sd, quantile(level / 2), median, quantile(1 - level / 2)  (type-7 / numpy "linear"),
 * split-R-hat (BDA3; posterior::rhat_basic), ESS (posterior::ess_basic: Geyer initial positive + monotone sequence),
 * pip (share of non-zero entries).  NaN entries (SAVS 0/0) are skipped by mean / sd / pip; a column holding any has NaN
 * order statistics and diagnostics. */
int bvhar_cuda_fit_record_summary(bvhar_fit* fit, const char* name, double level, double* out, int64_t cols);

/* Current sampler state of one unit, by member name of McmcTriangular (triangular.h:188-224) and
 * the prior classes: coef_mat, contem_coef, chol_lower, latent_innov, sqrt_sv, prior_alpha_prec,
 * prior_chol_prec, sparse_coef, sparse_contem, diag_vec, lvol_draw, lvol_init, lvol_sig, local_lev,
 * group_lev, global_lev, ...  Column-major, `len` doubles.  Used by the parity tests. */
int bvhar_cuda_fit_state_len(bvhar_fit* fit, const char* name, int64_t* len);
int bvhar_cuda_fit_get_state(bvhar_fit* fit, int window, int chain, const char* name, double* out,
                             int64_t len);
int bvhar_cuda_fit_set_state(bvhar_fit* fit, int window, int chain, const char* name,
                             const double* in, int64_t len);
/* Run a single updater of the sweep (1..9 in the order of triangular.h:102-115) on every unit,
 * with the Philox iteration counter set to `iter`.  Test/diagnostic entry. */
int bvhar_cuda_fit_run_stage(bvhar_fit* fit, int stage, int iter);

void bvhar_cuda_fit_destroy(bvhar_fit* fit);

/* ---- density forecast from stored draws (forecaster.h:27-267, 494-568) ------------------ */
typedef struct {
  int32_t abi_version;
  int32_t device;
  int32_t n_units;  /* independent forecasters (chains, or windows x chains) */
  int32_t dim;      /* k */
  int32_t nrow_coef;/* rows of the coefficient matrix without the constant: k*p or 3k */
  int32_t include_mean;
  int32_t cov_type; /* LDLT: d_record; SV: h (last time point) + sigh */
  int32_t step;     /* forecast horizon */
  int32_t var_lag;  /* p, or `month` for VHAR */
  int32_t is_vhar;  /* use har_trans (forecaster.h:338-340) */
  int32_t sv;       /* SvForecaster `sv` flag (forecaster.h:248-254) */
  int32_t get_lpl;
  int32_t n_draws;  /* draws per unit (after thinning / stability filter) */
  int32_t unit_id_base; /* Philox key word 1 of forecaster u is unit_id_base + u */
  const double* har_trans; /* (3k+1) x (month*k+1) column-major when is_vhar, else NULL */
  /* per unit, column-major n_draws x cols, concatenated over units */
  const double* coef_record;   /* n_draws x (nrow_coef*k + k*include_mean) */
  const double* contem_record; /* n_draws x q */
  const double* diag_record;   /* LDLT: n_draws x k (d_record) ; SV: n_draws x k (h at last time point) */
  const double* sigh_record;   /* SV: n_draws x k */
  const double* last_obs;      /* per unit: var_lag x k column-major, the last var_lag rows of the response */
  const double* valid_vec;     /* k (forecaster.h:802) or NULL */
  const uint32_t* seed;        /* n_units (seed_forecast, forecaster.h:1028) */
  /* level > 0 selects McmcVarSelectForecaster / McmcVharSelectForecaster (forecaster.h:348-416, factory :437-487): each
   * unit's coefficient draws are masked by its activity graph, RegRecords::computeActivity(level) (config.h:747-758):
   * entry i of coef_record is inactive (0) when the credible interval [tail_quantile<left>(level / 2),
   * tail_quantile<right>(1 - level / 2)] of its draws has end points of opposite sign, otherwise 1.1 (sic); the vector
   * is reshaped column-major to (num_coef / dim) x dim (forecaster.h:356).  0 = plain forecasters. */
  double level;
} bvhar_forecast_desc;

/* out_density: per unit, step x (n_draws*k) column-major (McmcForecaster::predictive_distn,
 * forecaster.h:39); out_lpl: per unit, `step` averages (forecaster.h:83) or NULL. */
int bvhar_cuda_forecast_density(const bvhar_forecast_desc* desc, double* out_density, double* out_lpl);

/* ---- rolling / expanding out-of-sample forecasts in one call ---------------------------------
 * Replaces McmcVarforecastRun / McmcVharforecastRun<McmcRollforecastRun | McmcExpandforecastRun, ...>::returnForecast()
 * (forecaster.h:575-1121; bound by src/triangular.cpp:159-630 roll_* / expand_* and _cta.cpp:40-208) for the windows
 * of `fit`'s window table: every (window, chain) is refitted (forecaster.h:755-790) and forecast from its own draws
 * (forecaster.h:798-806) without moving the draws to the host.  Window 0 of the reference reuses an existing fit
 * (forecaster.h:799-801): the caller passes windows >= 1 here (unit_id_base = n_chains) and forecasts window 0 with
 * bvhar_cuda_forecast_density. */
typedef struct {
  int32_t abi_version;
  int32_t step;     /* forecast horizon; only the last step is returned (forecaster.h:803) */
  int32_t var_lag;  /* p, or `month` for VHAR */
  int32_t is_vhar;
  int32_t sv;       /* SvForecaster `sv` flag */
  int32_t get_lpl;
  int32_t sparse;   /* forecast from the SAVS draws (alpha_sparse / c_sparse / a_sparse) */
  int32_t reserved;
  const double* har_trans;      /* (3k+1) x (month*k+1) column-major when is_vhar */
  const double* y_full;         /* y_rows x k column-major: training + test series the window table was built on */
  int64_t y_rows;
  const double* valid_vec;      /* k (y_test.row(step), forecaster.h:802) or NULL */
  const uint32_t* seed_forecast;/* n_chains (forecaster.h:1028) */
  double level;                 /* > 0: Select forecasters for every refitted window (forecaster.h:1024-1030, 1109-1115) */
} bvhar_outforecast_params;
/* out_forecast: [window][chain][draw * k + variable] (n_windows * n_chains * S * k doubles, S = kept draws);
 * out_lpl: [window][chain] average log-predictive likelihood over the steps, or NULL. */
int bvhar_cuda_outforecast_run(const bvhar_fit_desc* fit, const bvhar_outforecast_params* fc, double* out_forecast,
                               double* out_lpl);

/* ---- stability filter of the forecasters ---------------------------------------------------------
 * Replaces subsetStable (config.h:884-914, 1003-1034) -> is_stable / build_companion / root_unitcircle (draw.h:1265-1295)
 * for a batch of draws: out[i] = 1 when every eigenvalue of the VAR(1) companion matrix of draw i has modulus <
 * `threshold` (the forecasters pass 1, forecaster.h:283, 319).  coef: draw i at coef + i * ld, the first nrow * dim
 * entries being coef_record.row(i).head(num_alpha) (equation-major: alpha[j * nrow + r] = A(r, j)).  VAR: nrow = dim * lag,
 * har_trans = NULL.  VHAR: nrow = 3 * dim, lag = month and har_trans = the top-left 3 dim x (month dim) corner of the
 * HAR transformation, column-major with leading dimension ldh (forecaster.h:319).  The decision is taken on the device by
 * repeated squaring of companion / threshold (||B^N|| < 1 <=> stable; |trace(B^N)| > m => unstable); see
 * bvhar_b200/csrc/kernels_stable.cu. */
int bvhar_cuda_is_stable(int device, int64_t n_draws, int32_t dim, int32_t nrow, int32_t lag, const double* coef, int64_t ld,
                         const double* har_trans, int64_t ldh, double threshold, int32_t* out);

/* ---- spillover densities from stored draws ------------------------------------------------------
 * Replaces McmcSpilloverRun<RecordType>::returnSpillover() / McmcSpillover::computeSpillover (spillover.h:48-81, 246-260;
 * bound by src/triangular.cpp:636-647 and python/src/bvhar/_src/_cta.cpp) and the structural.h helpers it calls
 * (convert_var_to_vma :8-36, convert_vhar_to_vma :54-78, compute_vma_fevd :100-127, compute_sp_index / _to / _from /
 * _tot / _net :129-152) for a batch of draws.  Per draw i: coef + i * ld_coef holds alpha (equation-major, constants not
 * used), contem + i * ld_contem the q contemporaneous coefficients, sv + i * ld_sv the record the innovation scale D^(1/2)
 * comes from (updateDiag, config.h:876-881, 980-996): sv_kind 0 = d_record row (sqrt), 1 = log-volatilities at one time
 * point (exp(h / 2)), 2 = the whole h_record row, n_time x dim time-major (time average of exp(h / 2)), 3 = the whole
 * h_record row, one spillover PER TIME POINT (DynamicSvSpillover, spillover.h:445-507): the outputs then hold
 * n_time * n_draws draws ordered [time][draw].
 * VHAR: nrow = 3 dim, lag = month, har_trans = HAR transformation without constant, 3 dim x (month dim) column-major. */
typedef struct {
  int32_t abi_version;
  int32_t device;
  int32_t dim;
  int32_t nrow;   /* dim * lag (VAR) or 3 * dim (VHAR) */
  int32_t lag;    /* p, or month */
  int32_t step;   /* lag_max of McmcSpillover: horizon of the FEVD */
  int32_t sv_kind;
  int32_t n_time; /* sv_kind 2 */
  int64_t n_draws;
  const double* coef;
  int64_t ld_coef;
  const double* contem;
  int64_t ld_contem;
  const double* sv;
  int64_t ld_sv;
  const double* har_trans; /* NULL for VAR */
  int64_t ldh;
} bvhar_spillover_desc;
/* connect, net_pairwise: dim x (n_draws * dim) column-major; to, from: n_draws * dim; tot: n_draws
 * (returnSpilloverDensity, spillover.h:72-81; "net" = to - from). */
int bvhar_cuda_spillover(const bvhar_spillover_desc* desc, double* connect, double* to, double* from, double* tot, double* net_pairwise);

/* Rolling-window spillover with a refit per window: replaces DynamicLdltSpillover::returnSpillover() (spillover.h:272-443;
 * bound by src/triangular.cpp:649-709 dynamic_bvarldlt_spillover / dynamic_bvharldlt_spillover) for the window table of
 * `fit` (window i = rows [i, i + window) of y, design built once on the whole series as for the rolling forecasts).  Every
 * (window, chain) is refitted and its spillover computed from the device-resident draws; the window table is processed in
 * batches sized to the free device memory.  Works for both covariance models (SV fits use the time average of exp(h / 2),
 * which needs BVHAR_REC_LVOL records and is memory-hungry; LDLT is the reference's use).
 * to, from: [window][chain][draw * dim + variable]; tot: [window][chain][draw]; net = to - from. */
typedef struct {
  int32_t abi_version;
  int32_t step;    /* FEVD horizon */
  int32_t var_lag; /* p, or month */
  int32_t sparse;  /* use the SAVS draws */
  const double* har_trans; /* VHAR: 3 dim x (month dim) column-major without constant, else NULL */
  int64_t ldh;
} bvhar_dynspill_params;
int bvhar_cuda_dynamic_spillover(const bvhar_fit_desc* fit, const bvhar_dynspill_params* p, double* to, double* from, double* tot);

/* ---- conjugate Minnesota samplers ----
 * One call replaces the per-chain iteration loops of
 *   estimate_mniw    (src/mniw.cpp:130-190): McmcMniw::doPosteriorDraws, bayes/mniw/minnesota.h:241-278
 *   estimate_bvar_mh (src/mniw.cpp:46-110):  MhMinnesota::doPosteriorDraws, minnesota.h:413-510 (mh = 1)
 * i.e. sim_mn_iw (math/random.h:177-211) per iteration and, for mh = 1, the random-walk Metropolis step on (lambda, psi)
 * (minnesota.h:446-461, jointdens_hyperparam draw.h:59-72).  The posterior moments (Minnesota::estimateCoef /
 * estimateCov, minnesota.h:176-187) are inputs.  Every (chain, kept iteration) is an independent work item on the device
 * (addressed Philox stream: key (seed_chain[c], unit_id_base + c), iteration = mcmc_step).
 * Records follow thin_record (draw.h:1297-1310): rows = ceil((num_iter - num_burn) / thin), mcmc_step
 * num_burn + 1 + r * thin; every record is column-major rows x cols per chain, chains stacked. */
typedef struct {
  int32_t device;
  int32_t num_chains, num_iter, num_burn, thin;
  int32_t dim, dim_design;
  const double* mn_mean;  /* dim_design x dim, column-major: MinnFit::_coef */
  const double* mn_prec;  /* dim_design x dim_design:        MinnFit::_prec */
  const double* iw_scale; /* dim x dim:                      MinnFit::_iw_scale */
  double iw_shape;        /*                                 MinnFit::_iw_shape */
  const uint32_t* seed_chain; /* [num_chains] */
  uint32_t unit_id_base;
  int32_t mh; /* 0: McmcMniw; 1: MhMinnesota (the fields below are read) */
  double gam_shape, gam_rate;       /* param_prior$lambda$param (MhMinnSpec, minnesota.h:97-99); the shape <-> rate swap of
                                       minnesota.h:426 is applied inside, as the reference does */
  double invgam_shape, invgam_scl;  /* param_prior$sigma$param */
  const double* init_par;            /* [num_chains][1 + dim]: MhMinnInits par = (lambda, psi) */
  const double* init_hess;           /* [num_chains] (1 + dim) x (1 + dim) column-major: MhMinnInits hessian */
  const double* init_scale_variance; /* [num_chains]: MhMinnInits scale_variance */
} bvhar_mniw_desc;
/* alpha_record [chains][rows x dim_design*dim], sigma_record [chains][rows x dim*dim]; mh = 1 also fills
 * lambda_record [chains][rows], psi_record [chains][rows x dim], accept_record [chains][rows] (0 / 1); NULL skips one.
 * BVHAR_ERR_NUMERICAL: a proposed psi was negative (invgamma_dens STOPs, common.h:503) or a factorisation failed. */
int bvhar_cuda_mniw_run(const bvhar_mniw_desc* d, double* alpha_record, double* sigma_record, double* lambda_record,
                        double* psi_record, uint8_t* accept_record);

/* ---- fp64 building blocks exported for stage-level parity tests and the roofline bench ---- */
/* C[M x N] (row-major, ldc) = alpha * A^T B + beta * C with A stored K x M row-major (lda) and
 * B stored K x N row-major (ldb): the DMMA contraction that builds the SV-weighted Gram
 * X' diag(w) X for every (unit, equation) at once.  Host pointers; copies in and out. */
int bvhar_cuda_dgemm_tn(int device, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double* C, int64_t ldc);
/* Batched precision draw of draw.h:372-383: for each of `batch` systems, chol(P) with P packed
 * lower (d(d+1)/2, row-major packed), mean = P^-1 b, out = mean + chol(P)^-T z. */
int bvhar_cuda_draw_coef_batched(int device, int64_t batch, int32_t d, const double* P_packed,
                                 const double* b, const double* z, double* out);

/* ---- device-resident micro-benchmarks (roofline denominators; bench.py) ---------------------- */
/* mode 0: fp64 FMA issue peak, mode 1: fp64 DMMA (mma.sync.m8n8k4) issue peak, in TFLOP/s. */
int bvhar_cuda_bench_fp64_peak(int device, int mode, double* tflops);
/* cycles per DEPENDENT fp64 operation, one warp: kind 0 DFMA, 1 DMMA m8n8k4 accumulator chain, 2 rsqrt, 3 reciprocal, 4 DFMA through a
 * 64-bit shuffle.  Diagnostic (profiles/r02_fp64_peaks.md): the latency-bound kernels of the sweep are chains of these. */
int bvhar_cuda_bench_fp64_latency(int device, int kind, double* cycles);
/* The TN contraction kernel alone on synthetic device-resident operands, CUDA-event timed. */
int bvhar_cuda_bench_dgemm_tn(int device, int64_t M, int64_t N, int64_t K, int iters, double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* BVHAR_CUDA_H */
