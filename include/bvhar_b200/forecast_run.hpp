// forecast_run.hpp — header-only C++17 host adapter for the density forecast from stored draws.
//
//   bvhar_b200::McmcForecastRun<isSv>   <-  bvhar::McmcForecastRun<RegForecaster | SvForecaster>
//                                            inst/include/bvhar/src/bayes/triangular/forecaster.h:494-568
//       the same three constructors (VAR: lag; VHAR: week + month, or month + har_trans) with the reference's argument
//       order, and `returnForecast()` -> one step x (draws * dim) matrix per chain (predictive_distn, forecaster.h:39).
//       `fit_record` is what McmcRun::returnRecords() returned (one record map per chain); `stable = true` is the
//       subsetStable filter (config.h:884-914, 1003-1034), decided on the device by bvhar_cuda_is_stable; `level > 0`
//       selects the Select forecasters (forecaster.h:348-416): activity graph on the device, mask in the forecast kernel.
//
// As in mcmc_run.hpp, `Matrix` stands for Eigen::MatrixXd and `Records` for the per-chain LIST.  Errors are
// std::runtime_error with the reference's messages (forecaster.h:286, 322, 446).  No CPU fallback.
#pragma once
#include "mcmc_run.hpp"

namespace bvhar_b200 {

// build_vhar (math/design.h:35-58): (3 dim [+1]) x (month dim [+1]) HAR transformation, column-major.
inline Matrix build_vhar(int dim, int week, int month, bool include_mean) {
  const int64_t r = 3 * dim + (include_mean ? 1 : 0), c = (int64_t)month * dim + (include_mean ? 1 : 0);
  Matrix out(r, c);
  for (int i = 0; i < dim; ++i) {
    out(i, i) = 1.0;
    for (int l = 0; l < week; ++l) out(dim + i, (int64_t)l * dim + i) = 1.0 / week;
    for (int l = 0; l < month; ++l) out(2 * dim + i, (int64_t)l * dim + i) = 1.0 / month;
  }
  if (include_mean) out(3 * dim, (int64_t)month * dim) = 1.0;
  return out;
}

template <bool isSv>
class McmcForecastRun {
 public:
  // forecaster.h:497-510 (VAR)
  McmcForecastRun(int num_chains, int lag, int step, const Matrix& response_mat, bool sparse, double level,
                  const std::vector<Records>& fit_record, const std::vector<uint32_t>& seed_chain, bool include_mean, bool stable,
                  int /*nthreads*/, bool sv = true, int device = 0) {
    init(num_chains, lag, step, response_mat, nullptr, sparse, level, fit_record, seed_chain, include_mean, stable, sv, device);
  }
  // forecaster.h:511-525 (VHAR from week / month)
  McmcForecastRun(int num_chains, int week, int month, int step, const Matrix& response_mat, bool sparse, double level,
                  const std::vector<Records>& fit_record, const std::vector<uint32_t>& seed_chain, bool include_mean, bool stable,
                  int /*nthreads*/, bool sv = true, int device = 0) {
    const Matrix har = build_vhar((int)response_mat.cols, week, month, include_mean);
    init(num_chains, month, step, response_mat, &har, sparse, level, fit_record, seed_chain, include_mean, stable, sv, device);
  }
  // forecaster.h:526-539 (VHAR with a given transformation)
  McmcForecastRun(int num_chains, int month, int step, const Matrix& response_mat, const Matrix& har_trans, bool sparse, double level,
                  const std::vector<Records>& fit_record, const std::vector<uint32_t>& seed_chain, bool include_mean, bool stable,
                  int /*nthreads*/, bool sv = true, int device = 0) {
    init(num_chains, month, step, response_mat, &har_trans, sparse, level, fit_record, seed_chain, include_mean, stable, sv, device);
  }

  // forecaster.h:557-560
  std::vector<Matrix> returnForecast() { return forecastDensity(nullptr, nullptr, 0); }

  // McmcForecaster::forecastDensity(valid_vec) for every chain (forecaster.h:73-85): `lpl`, when given, receives the
  // chains x step log-predictive-likelihood averages (forecaster.h:83); `unit_id_base` = Philox key word 1 of chain 0
  // (the out-of-sample runners number their forecasters by (window, chain) unit).
  std::vector<Matrix> forecastDensity(const std::vector<double>* valid_vec, std::vector<double>* lpl, int unit_id_base) {
    bvhar_forecast_desc D = {};
    D.abi_version = BVHAR_CUDA_ABI_VERSION; D.device = device_; D.n_units = chains_; D.dim = dim_; D.nrow_coef = nrow_;
    D.include_mean = mean_; D.cov_type = isSv ? BVHAR_COV_SV : BVHAR_COV_LDLT; D.step = step_; D.var_lag = lag_;
    D.is_vhar = har_.data.empty() ? 0 : 1; D.sv = sv_; D.n_draws = draws_; D.unit_id_base = unit_id_base;
    D.get_lpl = valid_vec && lpl ? 1 : 0; D.valid_vec = D.get_lpl ? valid_vec->data() : nullptr;
    if (D.get_lpl && (int)valid_vec->size() != dim_) throw std::runtime_error("valid_vec needs one entry per variable");
    D.har_trans = har_.data.empty() ? nullptr : har_.data.data();
    D.coef_record = coef_.data(); D.contem_record = contem_.data(); D.diag_record = diag_.data();
    D.sigh_record = isSv ? sigh_.data() : nullptr; D.last_obs = last_.data(); D.seed = seeds_.data();
    D.level = level_;  // activity graph of every chain computed on the device (config.h:747-758)
    std::vector<double> out((size_t)chains_ * step_ * draws_ * dim_);
    if (D.get_lpl) lpl->assign((size_t)chains_ * step_, 0.0);
    detail::check(bvhar_cuda_forecast_density(&D, out.data(), D.get_lpl ? lpl->data() : nullptr));
    std::vector<Matrix> res;
    const size_t per = (size_t)step_ * draws_ * dim_;
    for (int c = 0; c < chains_; ++c) {
      Matrix m(step_, (int64_t)draws_ * dim_);
      std::copy(out.begin() + c * per, out.begin() + (c + 1) * per, m.data.begin());
      res.push_back(std::move(m));
    }
    return res;
  }

 private:
  static const Matrix& rec(const Records& r, const std::string& name) {
    auto it = r.find(name);
    if (it == r.end()) throw std::runtime_error("fit_record lacks '" + name + "'");
    return it->second;
  }
  void init(int num_chains, int ord, int step, const Matrix& response_mat, const Matrix* har, bool sparse, double level,
            const std::vector<Records>& fit_record, const std::vector<uint32_t>& seed_chain, bool include_mean, bool stable, bool sv,
            int device) {
    if (sparse && level > 0) throw std::runtime_error("If 'level > 0', 'spare' should be false.");  // forecaster.h:445-447
    level_ = level > 0 ? level : 0.0;  // level > 0: McmcVarSelectForecaster / McmcVharSelectForecaster (forecaster.h:348-416, 437-487)
    if ((int)fit_record.size() != num_chains || (int)seed_chain.size() != num_chains) throw std::runtime_error("fit_record / seed_chain need one entry per chain");
    if (response_mat.rows < ord) throw std::runtime_error("response_mat has fewer rows than the lag");
    chains_ = num_chains; lag_ = ord; step_ = step; dim_ = (int)response_mat.cols; mean_ = include_mean; sv_ = sv; device_ = device;
    seeds_ = seed_chain;
    if (har) har_ = *har;
    const std::string sfx = sparse ? "_sparse_record" : "_record";
    const int k = dim_, q = k * (k - 1) / 2;
    std::vector<std::vector<int64_t>> keep(num_chains);
    for (int c = 0; c < num_chains; ++c) {
      const Matrix& a = rec(fit_record[c], "alpha" + sfx);
      nrow_ = (int)(a.cols / k);
      if (har ? (nrow_ != 3 * k) : (nrow_ != k * ord)) throw std::runtime_error("coefficient record does not match the lag structure");
      std::vector<int32_t> ok((size_t)a.rows, 1);
      if (stable && a.rows > 0) {  // subsetStable: rows of the record, row-major for the device filter
        std::vector<double> rows((size_t)a.rows * a.cols);
        for (int64_t i = 0; i < a.rows; ++i)
          for (int64_t j = 0; j < a.cols; ++j) rows[(size_t)i * a.cols + j] = a(i, j);
        std::vector<double> hcorner;
        if (har) {
          hcorner.resize((size_t)3 * k * ord * k);
          for (int64_t cc = 0; cc < (int64_t)ord * k; ++cc)
            for (int r = 0; r < 3 * k; ++r) hcorner[(size_t)cc * 3 * k + r] = (*har)(r, cc);
        }
        detail::check(bvhar_cuda_is_stable(device, a.rows, k, nrow_, ord, rows.data(), a.cols, har ? hcorner.data() : nullptr, 3 * k, 1.0, ok.data()));
      }
      for (int64_t i = 0; i < a.rows; ++i)
        if (ok[(size_t)i]) keep[c].push_back(i);
    }
    draws_ = (int)keep[0].size();
    for (auto& kc : keep)
      if ((int)kc.size() != draws_) throw std::runtime_error("all chains must keep the same number of draws");
    if (draws_ == 0) throw std::runtime_error("No stable MCMC draws");  // forecaster.h:286, 322
    const int ccols = nrow_ * k + (include_mean ? k : 0);
    auto gather = [&](std::vector<double>& dst, int cols, auto&& value) {  // per chain: draws_ x cols column-major
      dst.assign((size_t)num_chains * draws_ * std::max(cols, 0), 0.0);
      for (int c = 0; c < num_chains; ++c)
        for (int j = 0; j < cols; ++j)
          for (int i = 0; i < draws_; ++i) dst[((size_t)c * cols + j) * draws_ + i] = value(c, keep[c][(size_t)i], j);
    };
    gather(coef_, ccols, [&](int c, int64_t i, int j) {
      return j < nrow_ * k ? rec(fit_record[c], "alpha" + sfx)(i, j) : rec(fit_record[c], "c" + sfx)(i, j - nrow_ * k);
    });
    gather(contem_, q, [&](int c, int64_t i, int j) { return rec(fit_record[c], "a" + sfx)(i, j); });
    if (isSv) {
      gather(diag_, k, [&](int c, int64_t i, int j) {  // h at the last time point (config.h:998-1001)
        const Records& r = fit_record[c];
        auto hl = r.find("h_last_record");
        if (hl != r.end()) return hl->second(i, j);
        const Matrix& h = rec(r, "h_record");
        return h(i, h.cols - k + j);
      });
      gather(sigh_, k, [&](int c, int64_t i, int j) { return rec(fit_record[c], "sigh_record")(i, j); });
    } else {
      gather(diag_, k, [&](int c, int64_t i, int j) { return rec(fit_record[c], "d_record")(i, j); });
    }
    last_.assign((size_t)num_chains * ord * k, 0.0);  // per chain: the last `ord` rows of the response, column-major
    for (int c = 0; c < num_chains; ++c)
      for (int j = 0; j < k; ++j)
        for (int l = 0; l < ord; ++l) last_[((size_t)c * k + j) * ord + l] = response_mat(response_mat.rows - ord + l, j);
  }

  int draws() const { return draws_; }

 private:
  int chains_ = 0, lag_ = 0, step_ = 0, dim_ = 0, nrow_ = 0, draws_ = 0, device_ = 0;
  bool mean_ = true, sv_ = true;
  double level_ = 0.0;
  Matrix har_;
  std::vector<uint32_t> seeds_;
  std::vector<double> coef_, contem_, diag_, sigh_, last_;
};

using LdltForecast = McmcForecastRun<false>;  // python/src/bvhar/_src/_cta.cpp:40-74
using SvForecast = McmcForecastRun<true>;

}  // namespace bvhar_b200
