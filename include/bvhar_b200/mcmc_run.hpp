// mcmc_run.hpp — header-only C++17 host adapter over the C-ABI of include/bvhar_cuda.h.
//
// Mirrors the reference's runner interface for the Gibbs hot path so that the Rcpp / pybind11 glue of
// INTEGRATION.md is a type substitution away:
//
//   bvhar_b200::McmcRun<isSv, isGroup>   <-  bvhar::McmcRun<McmcSv|McmcReg, isGroup>
//                                             inst/include/bvhar/src/bayes/triangular/triangular.h:1192-1233
//       same constructor argument order (num_chains, num_iter, num_burn, thin, x, y, param_cov, param_prior,
//       param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat, include_mean, seed_chain,
//       display_progress, nthreads) and `returnRecords()` -> one {record name -> matrix} map per chain with the
//       names of config.h:639-648, 814-820, 865-867, 967-971.
//
// Eigen / Rcpp::List / py::dict are not available in this build image, so `Matrix` (column-major) stands for
// Eigen::MatrixXd and `ParamList` (key -> flat column-major values) for LIST; keys are the reference's own
// (config.h:71-74, 100-101, 124-143, 180, 243-248, 298-301, 327, 352, 371-575).  Errors become
// std::runtime_error (the reference's STOP, core/commondefs.h:18).  There is no CPU fallback.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../bvhar_cuda.h"

namespace bvhar_b200 {

struct Matrix {  // column-major, like Eigen::MatrixXd
  int64_t rows = 0, cols = 0;
  std::vector<double> data;
  Matrix() = default;
  Matrix(int64_t r, int64_t c) : rows(r), cols(c), data((size_t)r * c, 0.0) {}
  double& operator()(int64_t i, int64_t j) { return data[(size_t)j * rows + i]; }
  double operator()(int64_t i, int64_t j) const { return data[(size_t)j * rows + i]; }
};
using ParamList = std::map<std::string, std::vector<double>>;
using Records = std::map<std::string, Matrix>;

namespace detail {
inline const std::vector<double>& need(const ParamList& l, const char* key, size_t n, const char* what) {
  auto it = l.find(key);
  if (it == l.end()) throw std::runtime_error(std::string(what) + " lacks '" + key + "'");
  if (n && it->second.size() != n) throw std::runtime_error(std::string(what) + "['" + key + "'] has the wrong length");
  return it->second;
}
inline double scalar(const ParamList& l, const char* key, const char* what) { return need(l, key, 0, what).at(0); }
inline const double* opt(const ParamList& l, const char* key) {
  auto it = l.find(key);
  return it == l.end() ? nullptr : it->second.data();
}
inline void check(int rc) {
  if (rc != BVHAR_OK) throw std::runtime_error(bvhar_cuda_last_error());
}
}  // namespace detail

// Prior mean / diagonal precision of the Minnesota-type priors from the dummy observations
// (build_ydummy / build_xdummy math/design.h:60-98, MinnParams config.h:124-151; the sampler only uses
// diag(kron(diag(1/sigma), Xd'Xd)), triangular.h:438-439).  Xd'Xd is diagonal, ((l+1) sigma_i / lambda)^2, so the
// LLT solve of config.h:149-151 reduces to a division.
inline void minnesota_moments(const ParamList& prior, int dim, int nrow, bool hierarchical, std::vector<double>& mean,
                              std::vector<double>& prec) {
  const int lag = (int)detail::scalar(prior, "p", "param_prior");
  const auto& sigma = detail::need(prior, "sigma", dim, "param_prior");
  if (lag * dim != nrow) throw std::runtime_error("Minnesota prior: 'p' does not match the coefficient matrix");
  const double lam = hierarchical ? 1.0 : detail::scalar(prior, "lambda", "param_prior");  // config.h:182-208
  std::vector<double> zero(dim, 0.0);
  const double* lagcoef[3] = {zero.data(), zero.data(), zero.data()};
  if (prior.count("delta")) lagcoef[0] = detail::need(prior, "delta", dim, "param_prior").data();
  else {
    lagcoef[0] = detail::need(prior, "daily", dim, "param_prior").data();
    lagcoef[1] = detail::need(prior, "weekly", dim, "param_prior").data();
    lagcoef[2] = detail::need(prior, "monthly", dim, "param_prior").data();
  }
  mean.assign((size_t)nrow * dim, 0.0);
  prec.assign((size_t)nrow * dim, 0.0);
  for (int j = 0; j < dim; ++j)
    for (int l = 0; l < lag; ++l)
      for (int i = 0; i < dim; ++i) {
        const double xd = (l + 1) * (sigma[i] / lam);  // diagonal entry of Xd
        prec[(size_t)j * nrow + l * dim + i] = (1.0 / sigma[j]) * (xd * xd);
        if (i == j && l < 3 && (lag > 1 || l == 0)) mean[(size_t)j * nrow + l * dim + i] = (lagcoef[l][i] * sigma[i] / lam) / xd;
      }
}

namespace detail {
inline std::vector<std::string> record_names(bvhar_fit* fit) {
  char buf[4096];
  check(bvhar_cuda_fit_record_names(fit, buf, sizeof buf));
  std::vector<std::string> names;
  for (const char* p = buf; *p;) {
    const char* e = p;
    while (*e && *e != '\n') ++e;
    names.emplace_back(p, e);
    p = *e ? e + 1 : e;
  }
  return names;
}
}  // namespace detail

// Fills a bvhar_fit_desc from the reference's constructor arguments (the LISTs parsed by config.h:50-575) and owns the
// temporaries the descriptor points into.  One window covering every row of x by default; the out-of-sample runners
// overwrite the window table, the seeds and unit_id_base.  `lvol_rows` = rows of the chains' `lvol` init blocks (the first
// window's length); with lvol_from_init the blocks are not read at all.
struct FitDescBuilder {
  bvhar_fit_desc desc{};
  std::vector<double> mmean, mprec;
  std::vector<bvhar_chain_init> inits;
  int32_t row0 = 0, nr = 0;

  FitDescBuilder(const Matrix& x, const Matrix& y, const ParamList& param_cov, const ParamList& param_prior, const ParamList& param_intercept,
                 const std::vector<ParamList>& param_init, int prior_type, const std::vector<int32_t>& grp_id,
                 const std::vector<int32_t>& own_id, const std::vector<int32_t>& cross_id, const std::vector<int32_t>& grp_mat,
                 bool include_mean, bool is_sv, bool is_group, int num_chains, int num_iter, int num_burn, int thin, int device,
                 uint32_t record_mask = BVHAR_REC_ALL, int64_t lvol_rows = -1) {
    using namespace detail;
    if (x.rows != y.rows) throw std::runtime_error("x and y need the same number of rows");
    if ((int)param_init.size() != num_chains) throw std::runtime_error("param_init needs one entry per chain");
    const int dim = (int)y.cols, d = (int)x.cols, q = dim * (dim - 1) / 2;
    const int n_alpha = dim * d - (include_mean ? dim : 0), nrow = n_alpha / dim, G = (int)grp_id.size();
    if ((int)grp_mat.size() != n_alpha) throw std::runtime_error("grp_mat has the wrong size");
    bvhar_fit_desc& D = desc;
    D.abi_version = BVHAR_CUDA_ABI_VERSION; D.device = device;
    D.n_windows = 1; D.n_chains = num_chains; D.dim = dim; D.dim_design = d; D.include_mean = include_mean;
    D.cov_type = is_sv ? BVHAR_COV_SV : BVHAR_COV_LDLT; D.prior_type = prior_type; D.is_group = is_group;
    D.num_iter = num_iter; D.num_burn = num_burn; D.thin = thin; D.record_mask = record_mask;
    D.n_rows_total = x.rows; D.x = x.data.data(); D.y = y.data.data();
    nr = (int32_t)x.rows;
    D.win_row0 = &row0; D.win_nrow = &nr;
    D.grp_mat = grp_mat.data(); D.grp_id = grp_id.data(); D.n_grp = G;
    D.own_id = own_id.data(); D.n_own = (int)own_id.size(); D.cross_id = cross_id.data(); D.n_cross = (int)cross_id.size();
    D.sig_shape = need(param_cov, "shape", dim, "param_cov").data();
    D.sig_scale = need(param_cov, "scale", dim, "param_cov").data();
    if (is_sv) {
      D.sv_init_mean = need(param_cov, "initial_mean", dim, "param_cov").data();
      D.sv_init_prec = need(param_cov, "initial_prec", (size_t)dim * dim, "param_cov").data();
    }
    if (include_mean) {
      D.mean_non = need(param_intercept, "mean_non", dim, "param_intercept").data();
      D.sd_non = scalar(param_intercept, "sd_non", "param_intercept");
    }
    bvhar_prior_params& P = D.prior;
    switch (prior_type) {
      case BVHAR_PRIOR_MINNESOTA:
      case BVHAR_PRIOR_HIERMINN:
        minnesota_moments(param_prior, dim, nrow, prior_type == BVHAR_PRIOR_HIERMINN, mmean, mprec);
        P.minn_prior_mean = mmean.data(); P.minn_prior_prec = mprec.data();
        if (prior_type == BVHAR_PRIOR_HIERMINN) {
          P.hmn_shape = scalar(param_prior, "shape", "param_prior"); P.hmn_rate = scalar(param_prior, "rate", "param_prior");
          P.hmn_grid_size = (int)scalar(param_prior, "grid_size", "param_prior");
        }
        break;
      case BVHAR_PRIOR_SSVS:
        P.ssvs_coef_s1 = need(param_prior, "coef_s1", G, "param_prior").data();
        P.ssvs_coef_s2 = need(param_prior, "coef_s2", G, "param_prior").data();
        P.ssvs_chol_s1 = scalar(param_prior, "chol_s1", "param_prior"); P.ssvs_chol_s2 = scalar(param_prior, "chol_s2", "param_prior");
        P.ssvs_coef_slab_shape = scalar(param_prior, "coef_slab_shape", "param_prior");
        P.ssvs_coef_slab_scl = scalar(param_prior, "coef_slab_scl", "param_prior");
        P.ssvs_chol_slab_shape = scalar(param_prior, "chol_slab_shape", "param_prior");
        P.ssvs_chol_slab_scl = scalar(param_prior, "chol_slab_scl", "param_prior");
        P.ssvs_coef_grid = (int)scalar(param_prior, "coef_grid", "param_prior");
        P.ssvs_chol_grid = (int)scalar(param_prior, "chol_grid", "param_prior");
        break;
      case BVHAR_PRIOR_HORSESHOE: break;
      case BVHAR_PRIOR_NG:
        P.ng_shape_sd = scalar(param_prior, "shape_sd", "param_prior");
        P.ng_group_shape = scalar(param_prior, "group_shape", "param_prior"); P.ng_group_scale = scalar(param_prior, "group_scale", "param_prior");
        P.ng_global_shape = scalar(param_prior, "global_shape", "param_prior"); P.ng_global_scale = scalar(param_prior, "global_scale", "param_prior");
        P.ng_contem_global_shape = scalar(param_prior, "contem_global_shape", "param_prior");
        P.ng_contem_global_scale = scalar(param_prior, "contem_global_scale", "param_prior");
        break;
      case BVHAR_PRIOR_DL:
        P.dl_grid_size = (int)scalar(param_prior, "grid_size", "param_prior");
        P.dl_shape = scalar(param_prior, "shape", "param_prior"); P.dl_scale = scalar(param_prior, "scale", "param_prior");
        break;
      case BVHAR_PRIOR_GDP:
        P.gdp_grid_shape = (int)scalar(param_prior, "grid_shape", "param_prior");
        P.gdp_grid_rate = (int)scalar(param_prior, "grid_rate", "param_prior");
        break;
      default: throw std::runtime_error("prior_type must be 1..7");
    }
    inits.resize(num_chains);
    for (int c = 0; c < num_chains; ++c) {
      const ParamList& in = param_init[c];
      bvhar_chain_init& I = inits[c];
      I = bvhar_chain_init{};
      I.init_coef = need(in, "init_coef", (size_t)d * dim, "param_init").data();
      I.init_contem = need(in, "init_contem", q, "param_init").data();
      if (is_sv) {
        I.lvol_init = need(in, "lvol_init", dim, "param_init").data();
        I.lvol = lvol_rows == 0 ? opt(in, "lvol") : need(in, "lvol", (size_t)(lvol_rows < 0 ? x.rows : lvol_rows) * dim, "param_init").data();
        I.lvol_sig = need(in, "lvol_sig", dim, "param_init").data();
      } else {
        I.init_diag = need(in, "init_diag", dim, "param_init").data();
      }
      auto sc = [&](const char* k2) { auto it = in.find(k2); return it == in.end() || it->second.empty() ? 0.0 : it->second[0]; };
      I.own_lambda = sc("own_lambda"); I.cross_lambda = sc("cross_lambda"); I.contem_lambda = sc("contem_lambda");
      I.coef_dummy = opt(in, "init_coef_dummy"); I.coef_mixture = opt(in, "coef_mixture"); I.chol_mixture = opt(in, "chol_mixture");
      I.coef_slab = opt(in, "coef_slab"); I.contem_slab = opt(in, "contem_slab");
      I.coef_spike_scl = sc("coef_spike_scl"); I.chol_spike_scl = sc("chol_spike_scl");
      I.local_sparsity = opt(in, "local_sparsity"); I.global_sparsity = sc("global_sparsity");
      I.group_sparsity = opt(in, "group_sparsity"); I.contem_local_sparsity = opt(in, "contem_local_sparsity");
      I.contem_global_sparsity = sc("contem_global_sparsity");
      I.local_shape = opt(in, "local_shape"); I.contem_shape = sc("contem_shape");
      I.group_rate = opt(in, "group_rate"); I.contem_rate = opt(in, "contem_rate");
      I.gamma_shape = sc("gamma_shape"); I.gamma_rate = sc("gamma_rate");
      I.contem_gamma_shape = sc("contem_gamma_shape"); I.contem_gamma_rate = sc("contem_gamma_rate");
    }
    D.init = inits.data();
  }
  FitDescBuilder(const FitDescBuilder&) = delete;
  FitDescBuilder& operator=(const FitDescBuilder&) = delete;
};

template <bool isSv, bool isGroup = true>
class McmcRun {
 public:
  McmcRun(int num_chains, int num_iter, int num_burn, int thin, const Matrix& x, const Matrix& y, const ParamList& param_cov,
          const ParamList& param_prior, const ParamList& param_intercept, const std::vector<ParamList>& param_init,
          int prior_type, const std::vector<int32_t>& grp_id, const std::vector<int32_t>& own_id,
          const std::vector<int32_t>& cross_id, const std::vector<int32_t>& grp_mat /* column-major (nrow x dim) */,
          bool include_mean, const std::vector<uint32_t>& seed_chain, bool display_progress = false, int nthreads = 1,
          int device = 0, uint32_t record_mask = BVHAR_REC_ALL)
      : num_chains_(num_chains), display_(display_progress) {
    (void)nthreads;  // all chains advance together on the GPU
    if ((param_cov.count("initial_mean") != 0) != isSv)
      throw std::runtime_error("param_cov does not match the sampler type ('initial_mean' selects SV, src/triangular.cpp:33)");
    if ((int)seed_chain.size() != num_chains) throw std::runtime_error("param_init / seed_chain need one entry per chain");
    FitDescBuilder B(x, y, param_cov, param_prior, param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat, include_mean,
                     isSv, isGroup, num_chains, num_iter, num_burn, thin, device, record_mask);
    B.desc.seed_chain = seed_chain.data();
    detail::check(bvhar_cuda_fit_create(&B.desc, &fit_));  // copies everything it needs: the builder may die now
  }
  McmcRun(const McmcRun&) = delete;
  McmcRun& operator=(const McmcRun&) = delete;
  ~McmcRun() {
    if (fit_) bvhar_cuda_fit_destroy(fit_);
  }

  // McmcRun::returnRecords (triangular.h:1230-1233): runs burn-in + sampling, returns the thinned records per chain.
  // Every record comes back with ONE bulk device-to-host copy (bvhar_cuda_fit_record_all: [chain][draw][column]) and is
  // then laid out column-major per chain, as Eigen / R / NumPy(order="F") hold it.
  std::vector<Records> returnRecords() {
    const int rc = bvhar_cuda_fit_run(fit_, display_ ? &McmcRun::log_progress : nullptr, nullptr);
    if (rc != BVHAR_OK && rc != BVHAR_ERR_NUMERICAL && rc != BVHAR_ERR_INTERRUPTED) detail::check(rc);
    const std::vector<std::string> names = detail::record_names(fit_);
    std::vector<Records> out(num_chains_);
    std::vector<double> bulk;
    for (const auto& nm : names) {
      int64_t r = 0, k = 0;
      detail::check(bvhar_cuda_fit_record_dims(fit_, nm.c_str(), &r, &k));
      bulk.resize((size_t)num_chains_ * r * k);
      if (r * k > 0) detail::check(bvhar_cuda_fit_record_all(fit_, nm.c_str(), bulk.data(), num_chains_, r, k));
      for (int c = 0; c < num_chains_; ++c) {
        Matrix m(r, k);
        const double* src = bulk.data() + (size_t)c * r * k;
        for (int64_t i = 0; i < r; ++i)
          for (int64_t j = 0; j < k; ++j) m(i, j) = src[(size_t)i * k + j];
        out[c].emplace(nm, std::move(m));
      }
    }
    return out;
  }

 private:
  static int log_progress(void*, int done, int total) {  // the 5 % messages of triangular.h:1255-1274
    std::fprintf(stderr, "[all chains] %d / %d\n", done, total);
    return 0;
  }
  bvhar_fit* fit_ = nullptr;
  int num_chains_;
  bool display_;
};

using McmcLdlt = McmcRun<false, true>;      // python/src/bvhar/_src/_cta.cpp:4-38
using McmcLdltGrp = McmcRun<false, false>;
using SvMcmc = McmcRun<true, true>;
using SvGrpMcmc = McmcRun<true, false>;

}  // namespace bvhar_b200
