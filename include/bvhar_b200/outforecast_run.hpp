// outforecast_run.hpp — header-only C++17 host adapter for the rolling / expanding out-of-sample forecast runners.
//
//   bvhar_b200::McmcVarforecastRun<Window, isSv, isGroup>    <-  bvhar::McmcVarforecastRun<McmcRollforecastRun | McmcExpandforecastRun,
//   bvhar_b200::McmcVharforecastRun<Window, isSv, isGroup>         RegForecaster | SvForecaster, isGroup> and McmcVharforecastRun<...>
//                                                                  inst/include/bvhar/src/bayes/triangular/forecaster.h:575-807 (base),
//                                                                  815-957 (roll / expand), 966-1121 (VAR / VHAR)
//       the reference's constructor argument order (forecaster.h:969-976, 1049-1056):
//         (y, lag | week, month, num_chains, num_iter, num_burn, thin, sparse, level, fit_record, param_reg, param_prior,
//          param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat, include_mean, stable, step, y_test,
//          get_lpl, seed_chain [num_horizon x num_chains], seed_forecast [num_chains], display_progress, nthreads, sv)
//       and `returnForecast()` -> { forecast[window][chain] : 1 x (draws * dim), lpl : num_horizon x num_chains }
//       (forecaster.h:641-650; only the last forecast step is kept, :803; lpl = mean over the steps, :804).
//
// What runs where: window 0 reuses the existing fit (`fit_record`, forecaster.h:799-801, 1014-1020) through
// McmcForecastRun; windows >= 1 are refitted AND forecast by ONE native call, bvhar_cuda_outforecast_run, which keeps the
// draws on the device (every (window, chain) unit sweeps concurrently; the window table is a set of row ranges of one
// design built on [y; y_test], forecaster.h:851-859, 920-928).  With `stable = true` the draws must pass the stability
// filter unit by unit (subsetStable may keep a different number of draws per unit), so the refit runs through McmcRun-style
// records and every unit is forecast by its own McmcForecastRun.  Same Philox keys on both paths: unit u = window * C + chain.
//
// As in mcmc_run.hpp, `Matrix` stands for Eigen::MatrixXd, `ParamList` for LIST, `Records` for the per-chain record LIST.
// Errors are std::runtime_error with the reference's messages.  No CPU fallback.
#pragma once
#include "forecast_run.hpp"

namespace bvhar_b200 {

enum class Window { Roll, Expand };  // McmcRollforecastRun (forecaster.h:815-877) | McmcExpandforecastRun (:884-957)

struct OutForecast {                          // the LIST of McmcOutforecastRun::returnForecast (forecaster.h:641-650)
  std::vector<std::vector<Matrix>> forecast;  // [window][chain]: 1 x (draws * dim)
  Matrix lpl;                                 // num_horizon x num_chains, filled when get_lpl
  bool has_lpl = false;
};

namespace detail {
// build_y0 / build_x0 (math/design.h:8-33): response rows [lag, T) and the lagged design [y_{t-1}, ..., y_{t-lag}, 1].
inline Matrix build_y0(const Matrix& y, int lag) {
  Matrix r(y.rows - lag, y.cols);
  for (int64_t j = 0; j < y.cols; ++j)
    for (int64_t i = 0; i < r.rows; ++i) r(i, j) = y(lag + i, j);
  return r;
}
inline Matrix build_x0(const Matrix& y, int lag, bool include_mean) {
  const int64_t n = y.rows - lag, k = y.cols;
  Matrix x(n, k * lag + (include_mean ? 1 : 0));
  for (int t = 0; t < lag; ++t)
    for (int64_t j = 0; j < k; ++j)
      for (int64_t i = 0; i < n; ++i) x(i, t * k + j) = y(lag - t - 1 + i, j);  // block t = build_y0(y, lag, lag - t)
  if (include_mean)
    for (int64_t i = 0; i < n; ++i) x(i, k * lag) = 1.0;
  return x;
}
inline Matrix times_transpose(const Matrix& a, const Matrix& b) {  // a * b'
  Matrix c(a.rows, b.rows);
  for (int64_t j = 0; j < b.rows; ++j)
    for (int64_t l = 0; l < a.cols; ++l) {
      const double v = b(j, l);
      if (v == 0.0) continue;
      for (int64_t i = 0; i < a.rows; ++i) c(i, j) += a(i, l) * v;
    }
  return c;
}
inline Matrix top_rows(const Matrix& m, int64_t n) {
  Matrix r(n, m.cols);
  for (int64_t j = 0; j < m.cols; ++j)
    for (int64_t i = 0; i < n; ++i) r(i, j) = m(i, j);
  return r;
}
}  // namespace detail

template <Window W, bool isSv, bool isGroup, bool isVhar>
class McmcOutforecastRun {
 public:
  OutForecast returnForecast() {
    OutForecast res;
    res.forecast.assign(num_horizon_, std::vector<Matrix>(num_chains_));
    res.has_lpl = get_lpl_;
    if (get_lpl_) res.lpl = Matrix(num_horizon_, num_chains_);
    const int k = dim_, C = num_chains_;
    const std::vector<double>* valid = valid_.empty() ? nullptr : &valid_;
    // ---- window 0: the existing fit (forecaster.h:799-801) ----
    {
      const Matrix y0 = detail::top_rows(tot_, lag_ + win_nrow_[0]);
      forecastUnits(fit_record_, y0, 0, res, 0, valid);
    }
    if (num_horizon_ == 1) return res;
    // ---- windows >= 1 ----
    const int Wn = num_horizon_ - 1;
    const Matrix x_tot = isVhar ? detail::times_transpose(detail::build_x0(tot_, lag_, include_mean_), har_) : detail::build_x0(tot_, lag_, include_mean_);
    const Matrix y_tot = detail::build_y0(tot_, lag_);
    std::vector<int32_t> row0(win_row0_.begin() + 1, win_row0_.end()), nrow(win_nrow_.begin() + 1, win_nrow_.end());
    std::vector<uint32_t> seeds((size_t)Wn * C);  // seed_chain(window, chain), forecaster.h:871
    for (int w = 1; w < num_horizon_; ++w)
      for (int c = 0; c < C; ++c) seeds[(size_t)(w - 1) * C + c] = seed_chain_[(size_t)w * C + c];
    FitDescBuilder B(x_tot, y_tot, param_reg_, param_prior_, param_intercept_, param_init_, prior_type_, grp_id_, own_id_, cross_id_, grp_mat_,
                     include_mean_, isSv, isGroup, C, num_iter_, num_burn_, thin_, device_, BVHAR_REC_ALL,
                     (isSv && W == Window::Roll) ? win_nrow_[0] : 0);  // rolling SV windows start from the chains' lvol blocks
    bvhar_fit_desc& D = B.desc;
    D.n_windows = Wn; D.win_row0 = row0.data(); D.win_nrow = nrow.data(); D.seed_chain = seeds.data();
    D.unit_id_base = C;  // unit u = window * C + chain, windows >= 1
    D.lvol_from_init = (isSv && W == Window::Expand) ? 1 : 0;  // SvInits(init, num_design), config.h:416-420, forecaster.h:936-944
    const int64_t S = (std::max(0, num_iter_ - num_burn_) + thin_ - 1) / thin_;
    if (!stable_) {
      bvhar_outforecast_params P = {};
      P.abi_version = BVHAR_CUDA_ABI_VERSION; P.step = step_; P.var_lag = lag_; P.is_vhar = isVhar; P.sv = sv_; P.get_lpl = get_lpl_ && valid;
      P.sparse = sparse_; P.level = level_;
      P.har_trans = isVhar ? har_.data.data() : nullptr; P.y_full = tot_.data.data(); P.y_rows = tot_.rows;
      P.valid_vec = valid ? valid_.data() : nullptr; P.seed_forecast = seed_forecast_.data();
      std::vector<double> out((size_t)Wn * C * S * k), lpl((size_t)Wn * C, 0.0);
      detail::check(bvhar_cuda_outforecast_run(&D, &P, out.data(), P.get_lpl ? lpl.data() : nullptr));
      for (int w = 1; w < num_horizon_; ++w)
        for (int c = 0; c < C; ++c) {
          Matrix m(1, S * k);
          const double* src = out.data() + ((size_t)(w - 1) * C + c) * S * k;
          std::copy(src, src + S * k, m.data.begin());
          res.forecast[w][c] = std::move(m);
          if (get_lpl_ && valid) res.lpl(w, c) = lpl[(size_t)(w - 1) * C + c];
        }
      return res;
    }
    // stable = true: refit, bring the draws back, filter + forecast every unit on its own
    D.record_mask = BVHAR_REC_COEF | BVHAR_REC_SPARSE | BVHAR_REC_LVOL_LAST;
    bvhar_fit* fit = nullptr;
    detail::check(bvhar_cuda_fit_create(&D, &fit));
    struct Guard { bvhar_fit* f; ~Guard() { if (f) bvhar_cuda_fit_destroy(f); } } guard{fit};
    const int rc = bvhar_cuda_fit_run(fit, nullptr, nullptr);
    if (rc != BVHAR_OK && rc != BVHAR_ERR_NUMERICAL) detail::check(rc);
    const std::vector<std::string> names = detail::record_names(fit);
    for (int w = 1; w < num_horizon_; ++w) {
      std::vector<Records> recs(C);
      for (int c = 0; c < C; ++c)
        for (const auto& nm : names) {
          int64_t r = 0, cc = 0;
          detail::check(bvhar_cuda_fit_record_dims(fit, nm.c_str(), &r, &cc));
          Matrix m(r, cc);
          if (r * cc > 0) detail::check(bvhar_cuda_fit_record(fit, w - 1, c, nm.c_str(), m.data.data(), r, cc));
          recs[c].emplace(nm, std::move(m));
        }
      const Matrix yw = detail::top_rows(tot_, lag_ + win_row0_[w] + win_nrow_[w]);
      forecastUnits(recs, yw, w, res, w * C, valid);
    }
    return res;
  }

 protected:
  McmcOutforecastRun(const Matrix& y, int lag, const Matrix* har, int num_chains, int num_iter, int num_burn, int thin, bool sparse,
                     double level, const std::vector<Records>& fit_record, const ParamList& param_reg, const ParamList& param_prior,
                     const ParamList& param_intercept, const std::vector<ParamList>& param_init, int prior_type,
                     const std::vector<int32_t>& grp_id, const std::vector<int32_t>& own_id, const std::vector<int32_t>& cross_id,
                     const std::vector<int32_t>& grp_mat, bool include_mean, bool stable, int step, const Matrix& y_test, bool get_lpl,
                     const std::vector<uint32_t>& seed_chain /* row-major num_horizon x num_chains */,
                     const std::vector<uint32_t>& seed_forecast, bool display_progress, int /*nthreads*/, bool sv, int device)
      : dim_((int)y.cols), lag_(lag), num_chains_(num_chains), num_iter_(num_iter), num_burn_(num_burn), thin_(thin < 1 ? 1 : thin),
        prior_type_(prior_type), step_(step), device_(device), include_mean_(include_mean), stable_(stable), sparse_(sparse),
        get_lpl_(get_lpl), sv_(sv), display_(display_progress), level_(level > 0 ? level : 0.0), fit_record_(fit_record),
        param_reg_(param_reg), param_prior_(param_prior), param_intercept_(param_intercept), param_init_(param_init), grp_id_(grp_id),
        own_id_(own_id), cross_id_(cross_id), grp_mat_(grp_mat), seed_chain_(seed_chain), seed_forecast_(seed_forecast) {
    if ((param_reg.count("initial_mean") != 0) != isSv)
      throw std::runtime_error("param_reg does not match the forecaster type ('initial_mean' selects SV, forecaster.h:936)");
    if (y_test.cols != y.cols) throw std::runtime_error("y_test needs the columns of y");
    num_horizon_ = (int)y_test.rows - step + 1;  // forecaster.h:586
    if (num_horizon_ < 1) throw std::runtime_error("y_test needs at least 'step' rows");
    if ((int)seed_chain.size() != num_horizon_ * num_chains || (int)seed_forecast.size() != num_chains)
      throw std::runtime_error("seed_chain must be num_horizon x num_chains and seed_forecast num_chains long");
    if (level_ > 0) sparse_ = false;  // forecaster.h:607-609
    if (har) har_ = *har;
    tot_ = Matrix(y.rows + y_test.rows, y.cols);  // tot_mat << y, y_test (forecaster.h:852-854)
    for (int64_t j = 0; j < y.cols; ++j) {
      for (int64_t i = 0; i < y.rows; ++i) tot_(i, j) = y(i, j);
      for (int64_t i = 0; i < y_test.rows; ++i) tot_(y.rows + i, j) = y_test(i, j);
    }
    const int n0 = (int)y.rows - lag;
    for (int w = 0; w < num_horizon_; ++w) {  // roll: rows w .. w + T - 1; expand: rows 0 .. T + w - 1 (forecaster.h:855-858, 924-927)
      win_row0_.push_back(W == Window::Roll ? w : 0);
      win_nrow_.push_back(W == Window::Roll ? n0 : n0 + w);
    }
    if ((int64_t)step < y_test.rows) {  // valid_vec = y_test.row(step): the same row for every window (forecaster.h:802)
      valid_.resize(dim_);
      for (int j = 0; j < dim_; ++j) valid_[j] = y_test(step, j);
    }
  }

 private:
  // One McmcForecastRun per unit (the filter may keep different numbers of draws), or one for all chains of the window.
  void forecastUnits(const std::vector<Records>& recs, const Matrix& y_window, int window, OutForecast& res, int unit_base,
                     const std::vector<double>* valid) {
    const int C = num_chains_;
    auto one = [&](const std::vector<Records>& r, const std::vector<uint32_t>& sd, int base, int first_chain) {
      std::vector<double> lpl;
      std::vector<Matrix> dens;
      if (isVhar) {
        McmcForecastRun<isSv> f((int)r.size(), lag_, step_, y_window, har_, sparse_, level_, r, sd, include_mean_, stable_, 1, sv_, device_);
        dens = f.forecastDensity(valid, get_lpl_ ? &lpl : nullptr, base);
      } else {
        McmcForecastRun<isSv> f((int)r.size(), lag_, step_, y_window, sparse_, level_, r, sd, include_mean_, stable_, 1, sv_, device_);
        dens = f.forecastDensity(valid, get_lpl_ ? &lpl : nullptr, base);
      }
      for (size_t c = 0; c < r.size(); ++c) {
        const Matrix& d = dens[c];
        Matrix last(1, d.cols);  // .bottomRows(1), forecaster.h:803
        for (int64_t j = 0; j < d.cols; ++j) last(0, j) = d(d.rows - 1, j);
        res.forecast[window][first_chain + c] = std::move(last);
        if (get_lpl_ && valid) {
          double s = 0.0;
          for (int h = 0; h < step_; ++h) s += lpl[c * step_ + h];
          res.lpl(window, first_chain + (int)c) = s / step_;  // returnLpl() = lpl.mean(), forecaster.h:104-106, 804
        }
      }
    };
    if ((int)recs.size() != C) throw std::runtime_error("fit_record needs one entry per chain");
    if (!stable_) return one(recs, seed_forecast_, unit_base, 0);
    for (int c = 0; c < C; ++c) one({recs[c]}, {seed_forecast_[c]}, unit_base + c, c);
  }

  int dim_, lag_, num_chains_, num_iter_, num_burn_, thin_, prior_type_, step_, device_, num_horizon_ = 0;
  bool include_mean_, stable_, sparse_, get_lpl_, sv_, display_;
  double level_;
  Matrix har_, tot_;
  std::vector<Records> fit_record_;
  ParamList param_reg_, param_prior_, param_intercept_;
  std::vector<ParamList> param_init_;
  std::vector<int32_t> grp_id_, own_id_, cross_id_, grp_mat_, win_row0_, win_nrow_;
  std::vector<uint32_t> seed_chain_, seed_forecast_;
  std::vector<double> valid_;
};

// forecaster.h:966-1037
template <Window W, bool isSv, bool isGroup = true>
class McmcVarforecastRun : public McmcOutforecastRun<W, isSv, isGroup, false> {
 public:
  McmcVarforecastRun(const Matrix& y, int lag, int num_chains, int num_iter, int num_burn, int thin, bool sparse, double level,
                     const std::vector<Records>& fit_record, const ParamList& param_reg, const ParamList& param_prior,
                     const ParamList& param_intercept, const std::vector<ParamList>& param_init, int prior_type,
                     const std::vector<int32_t>& grp_id, const std::vector<int32_t>& own_id, const std::vector<int32_t>& cross_id,
                     const std::vector<int32_t>& grp_mat, bool include_mean, bool stable, int step, const Matrix& y_test, bool get_lpl,
                     const std::vector<uint32_t>& seed_chain, const std::vector<uint32_t>& seed_forecast, bool display_progress,
                     int nthreads, bool sv = true, int device = 0)
      : McmcOutforecastRun<W, isSv, isGroup, false>(y, lag, nullptr, num_chains, num_iter, num_burn, thin, sparse, level, fit_record, param_reg,
                                                    param_prior, param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat,
                                                    include_mean, stable, step, y_test, get_lpl, seed_chain, seed_forecast, display_progress,
                                                    nthreads, sv, device) {}
};

// forecaster.h:1046-1121: lag = month, design = build_x0(., month) * har_trans' (:1105-1107)
template <Window W, bool isSv, bool isGroup = true>
class McmcVharforecastRun : public McmcOutforecastRun<W, isSv, isGroup, true> {
 public:
  McmcVharforecastRun(const Matrix& y, int week, int month, int num_chains, int num_iter, int num_burn, int thin, bool sparse, double level,
                      const std::vector<Records>& fit_record, const ParamList& param_reg, const ParamList& param_prior,
                      const ParamList& param_intercept, const std::vector<ParamList>& param_init, int prior_type,
                      const std::vector<int32_t>& grp_id, const std::vector<int32_t>& own_id, const std::vector<int32_t>& cross_id,
                      const std::vector<int32_t>& grp_mat, bool include_mean, bool stable, int step, const Matrix& y_test, bool get_lpl,
                      const std::vector<uint32_t>& seed_chain, const std::vector<uint32_t>& seed_forecast, bool display_progress,
                      int nthreads, bool sv = true, int device = 0)
      : McmcVharforecastRun(build_vhar((int)y.cols, week, month, include_mean), y, month, num_chains, num_iter, num_burn, thin, sparse, level,
                            fit_record, param_reg, param_prior, param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat,
                            include_mean, stable, step, y_test, get_lpl, seed_chain, seed_forecast, display_progress, nthreads, sv, device) {}

 private:
  McmcVharforecastRun(const Matrix& har, const Matrix& y, int month, int num_chains, int num_iter, int num_burn, int thin, bool sparse,
                      double level, const std::vector<Records>& fit_record, const ParamList& param_reg, const ParamList& param_prior,
                      const ParamList& param_intercept, const std::vector<ParamList>& param_init, int prior_type,
                      const std::vector<int32_t>& grp_id, const std::vector<int32_t>& own_id, const std::vector<int32_t>& cross_id,
                      const std::vector<int32_t>& grp_mat, bool include_mean, bool stable, int step, const Matrix& y_test, bool get_lpl,
                      const std::vector<uint32_t>& seed_chain, const std::vector<uint32_t>& seed_forecast, bool display_progress,
                      int nthreads, bool sv, int device)
      : McmcOutforecastRun<W, isSv, isGroup, true>(y, month, &har, num_chains, num_iter, num_burn, thin, sparse, level, fit_record, param_reg,
                                                   param_prior, param_intercept, param_init, prior_type, grp_id, own_id, cross_id, grp_mat,
                                                   include_mean, stable, step, y_test, get_lpl, seed_chain, seed_forecast, display_progress,
                                                   nthreads, sv, device) {}
};

// The pybind11 names of python/src/bvhar/_src/_cta.cpp:76-208
using LdltVarRoll = McmcVarforecastRun<Window::Roll, false, true>;
using LdltVarExpand = McmcVarforecastRun<Window::Expand, false, true>;
using LdltVharRoll = McmcVharforecastRun<Window::Roll, false, true>;
using LdltVharExpand = McmcVharforecastRun<Window::Expand, false, true>;
using LdltGrpVarRoll = McmcVarforecastRun<Window::Roll, false, false>;
using LdltGrpVarExpand = McmcVarforecastRun<Window::Expand, false, false>;
using LdltGrpVharRoll = McmcVharforecastRun<Window::Roll, false, false>;
using LdltGrpVharExpand = McmcVharforecastRun<Window::Expand, false, false>;
using SvVarRoll = McmcVarforecastRun<Window::Roll, true, true>;
using SvVarExpand = McmcVarforecastRun<Window::Expand, true, true>;
using SvVharRoll = McmcVharforecastRun<Window::Roll, true, true>;
using SvVharExpand = McmcVharforecastRun<Window::Expand, true, true>;
using SvGrpVarRoll = McmcVarforecastRun<Window::Roll, true, false>;
using SvGrpVarExpand = McmcVarforecastRun<Window::Expand, true, false>;
using SvGrpVharRoll = McmcVharforecastRun<Window::Roll, true, false>;
using SvGrpVharExpand = McmcVharforecastRun<Window::Expand, true, false>;

}  // namespace bvhar_b200
