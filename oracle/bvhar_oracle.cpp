// TEST INFRASTRUCTURE ONLY — CPU oracle for the bvhar Gibbs hot path (parity UNPINNED: the
// reference cannot be built in this image — no R/Rcpp/Eigen/Boost — and its own tests hold no
// numeric expectations for this path; see DESIGN.md "Oracle").  Nothing under oracle/ is
// imported, linked or executed by the product; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// A restatement of the reference's corrected-triangular Gibbs sampler, one function per
// reference function, each citing the file:line it follows.  Paths are relative to
// /root/reference/inst/include/bvhar/src:
//   tri.h  = bayes/triangular/triangular.h     cfg.h = bayes/triangular/config.h
//   fc.h   = bayes/triangular/forecaster.h     draw.h = bayes/misc/draw.h
//   design.h = math/design.h                   common.h = core/common.h
//
// Two formulations of the two expensive updaters, selected per fit:
//   faithful  — literally what the reference does: Kronecker design per equation (tri.h:286),
//               dense n x n precision per log-volatility series (draw.h:454-487);
//   efficient — the algebraically identical SV-weighted sufficient statistics
//               S_i = X' diag(1/s_i^2) X and the tridiagonal form of the SV precision; this is
//               the arithmetic the CUDA engine implements.
// tests/ check faithful == efficient to rounding, and the CUDA engine against `efficient`.
#include <omp.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include <string>

#include "../include/bvhar_cuda.h"
#include "oracle_linalg.h"
#include "oracle_rng.h"

namespace orc {

static const double KSC_P[7] = {0.0073, 0.10556, 0.00002, 0.04395, 0.34001, 0.24566, 0.2575};  // draw.h:445
static const double KSC_M[7] = {-10.12999, -3.97281, -8.56686, 2.77786, 0.61942, 1.79518, -1.08819};  // draw.h:447
static const double KSC_V[7] = {5.79596, 2.61369, 5.17950, 0.16735, 0.64009, 0.34023, 1.26261};  // draw.h:450

struct Shared {
  int W, C, k, d, N, nrow, q, G;
  bool mean, sv, ggl, faithful, philox;
  int prior;
  int num_iter, num_burn, thin;
  uint32_t rec_mask;
  bool lvol_from_init;
  std::vector<int> grp_vec, grp_id;
  std::set<int> own;
  Vec sig_shp, sig_scl, sv_mean, mean_non;
  Mat sv_prec;
  double sd_non;
  bvhar_prior_params pp;
  Vec minn_mean, minn_prec, ssvs_s1, ssvs_s2;
};

struct Window {
  int row0, n;
  Mat X, Y;
  Vec xnorm;  // column squared norms of X (draw.h:425)
  Mat XtX, XtY;
};

struct Records {
  int cap = 0, filled = 0;
  std::map<std::string, Mat> m;  // cap x cols, column-major; row s = posterior draw s (row 0 of the
                                 // reference's (iter+1)-row matrices is dropped on return, tri.h:141)
  std::vector<std::string> order;
  void add(const std::string& name, int cols) {
    m[name] = Mat(cap, cols);
    order.push_back(name);
  }
};

struct Chain {
  const Shared* S;
  const Window* Wd;
  Rng rng;
  int n, k, d, N, nrow, q, G;
  // McmcTriangular members (tri.h:188-224)
  Mat coef_mat, chol_lower, latent_innov, sqrt_sv, sparse_coef;
  Vec contem_coef, prior_alpha_mean, prior_alpha_prec, alpha_penalty, prior_chol_mean, prior_chol_prec;
  Vec sparse_contem, coef_vec;
  // McmcReg / McmcSv
  Vec diag_vec, lvol_init, lvol_sig;
  Mat lvol_draw;
  // prior state (union over the prior classes, tri.h:534-547, 629-642, 726-740, 824-835, 914-925, 999-1007)
  double own_lambda = 1, cross_lambda = 1, contem_lambda = 1;
  Vec coef_dummy, coef_weight, contem_dummy, contem_weight, coef_slab, contem_slab, slab_weight;
  double spike_scl = 0, contem_spike_scl = 0;
  Vec local_lev, group_lev, shrink_fac, latent_local, latent_group, coef_var;
  double global_lev = 1, latent_global = 0;
  Vec contem_local_lev, contem_global_lev, contem_var, latent_contem_local, latent_contem_global;
  Vec local_shape, local_shape_fac, contem_fac;
  double contem_shape = 0;
  double dir_concen = 0, contem_dir_concen = 0;
  Vec group_rate, group_rate_fac, contem_rate;
  double coef_gamma_shape = 0, coef_gamma_rate = 0, contem_gamma_shape = 0, contem_gamma_rate = 0;
  Records rec;
  uint32_t sweep_count = 0;
  int inner_threads = 1;

  // ---------------------------------------------------------------- construction
  // McmcTriangular ctor tri.h:40-71, McmcReg tri.h:363-367, McmcSv tri.h:388-396.
  void init(const Shared* S_, const Window* W_, const bvhar_chain_init& in, uint32_t seed, uint32_t unit) {
    S = S_; Wd = W_;
    n = Wd->n; k = S->k; d = S->d; N = S->N; nrow = S->nrow; q = S->q; G = S->G;
    rng.seed(seed, unit, S->philox);
    coef_mat = Mat(d, k);
    std::memcpy(coef_mat.data(), in.init_coef, sizeof(double) * d * k);
    contem_coef.assign(in.init_contem, in.init_contem + q);
    prior_alpha_mean.assign((size_t)k * d, 0.0);
    prior_alpha_prec.assign((size_t)k * d, 0.0);
    alpha_penalty.assign(N, 0.0);
    prior_chol_mean.assign(q, 0.0);
    prior_chol_prec.assign(q, 1.0);
    sparse_coef = Mat(d, k);
    sparse_contem.assign(q, 0.0);
    chol_lower = build_inv_lower(k, contem_coef.data());
    latent_innov = Mat(n, k);
    update_latent();
    sqrt_sv = Mat(n, k);
    if (S->mean) {
      for (int j = 0; j < k; ++j) {
        prior_alpha_mean[N + j] = S->mean_non[j];
        prior_alpha_prec[N + j] = 1.0 / (S->sd_non * S->sd_non);
      }
    }
    coef_vec.assign((size_t)k * d, 0.0);
    refresh_coef_vec();
    if (!S->sv) {
      diag_vec.assign(in.init_diag, in.init_diag + k);
    } else {
      lvol_init.assign(in.lvol_init, in.lvol_init + k);
      lvol_sig.assign(in.lvol_sig, in.lvol_sig + k);
      lvol_draw = Mat(n, k);
      if (S->lvol_from_init || in.lvol == nullptr) {  // SvInits(init, num_design) cfg.h:416-420
        for (int i = 0; i < k; ++i)
          for (int t = 0; t < n; ++t) lvol_draw(t, i) = lvol_init[i];
      } else {
        std::memcpy(lvol_draw.data(), in.lvol, sizeof(double) * n * k);
      }
    }
    init_prior(in);
    init_records();
  }

  void refresh_coef_vec() {  // tri.h:65-68, 300-301, 312
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < nrow; ++i) coef_vec[(size_t)j * nrow + i] = coef_mat(i, j);
    if (S->mean)
      for (int j = 0; j < k; ++j) coef_vec[N + j] = coef_mat(nrow, j);
  }

  void init_prior(const bvhar_chain_init& in) {
    const bvhar_prior_params& pp = S->pp;
    switch (S->prior) {
      case BVHAR_PRIOR_MINNESOTA:  // tri.h:437-443
        for (int i = 0; i < N; ++i) { prior_alpha_mean[i] = S->minn_mean[i]; prior_alpha_prec[i] = S->minn_prec[i]; }
        break;
      case BVHAR_PRIOR_HIERMINN:  // tri.h:474-494; cross_id is empty in the shipped code (draw.h:75)
        own_lambda = in.own_lambda; cross_lambda = in.cross_lambda; contem_lambda = in.contem_lambda;
        for (int i = 0; i < N; ++i) { prior_alpha_mean[i] = S->minn_mean[i]; prior_alpha_prec[i] = S->minn_prec[i] / own_lambda; }
        for (int i = 0; i < q; ++i) prior_chol_prec[i] /= contem_lambda;
        break;
      case BVHAR_PRIOR_SSVS:  // tri.h:563-577
        coef_dummy.assign(in.coef_dummy, in.coef_dummy + N);
        coef_weight.assign(in.coef_mixture, in.coef_mixture + G);
        contem_dummy.assign(q, 1.0);
        contem_weight.assign(in.chol_mixture, in.chol_mixture + q);
        coef_slab.assign(in.coef_slab, in.coef_slab + N);
        contem_slab.assign(in.contem_slab, in.contem_slab + q);
        spike_scl = in.coef_spike_scl;
        contem_spike_scl = in.coef_spike_scl;  // tri.h:569 initialises from the *coef* scale
        slab_weight.assign(N, 1.0);
        break;
      case BVHAR_PRIOR_HORSESHOE:  // tri.h:659-669
        local_lev.assign(in.local_sparsity, in.local_sparsity + N);
        group_lev.assign(in.group_sparsity, in.group_sparsity + G);
        global_lev = S->ggl ? in.global_sparsity : 1.0;
        shrink_fac.assign(N, 0.0); latent_local.assign(N, 0.0); latent_group.assign(G, 0.0); latent_global = 0.0;
        coef_var.assign(N, 0.0);
        contem_local_lev.assign(in.contem_local_sparsity, in.contem_local_sparsity + q);
        contem_global_lev.assign(1, in.contem_global_sparsity);
        contem_var.assign(q, 0.0); latent_contem_local.assign(q, 0.0); latent_contem_global.assign(1, 0.0);
        break;
      case BVHAR_PRIOR_NG:  // tri.h:757-770
        local_shape.assign(in.local_shape, in.local_shape + G);
        local_shape_fac.assign(N, 1.0);
        contem_shape = in.contem_shape;
        local_lev.assign(in.local_sparsity, in.local_sparsity + N);
        group_lev.assign(in.group_sparsity, in.group_sparsity + G);
        global_lev = S->ggl ? in.global_sparsity : 1.0;
        coef_var.assign(N, 0.0);
        contem_global_lev.assign(1, in.contem_global_sparsity);
        contem_fac.resize(q);
        for (int i = 0; i < q; ++i) contem_fac[i] = contem_global_lev[0] * in.contem_local_sparsity[i];
        break;
      case BVHAR_PRIOR_DL:  // tri.h:852-863
        local_lev.assign(in.local_sparsity, in.local_sparsity + N);
        group_lev.assign(G, 0.0);
        global_lev = S->ggl ? in.global_sparsity : 1.0;
        latent_local.assign(N, 0.0); coef_var.assign(N, 0.0);
        contem_local_lev.assign(in.contem_local_sparsity, in.contem_local_sparsity + q);
        contem_global_lev.assign(1, in.contem_global_sparsity);
        latent_contem_local.assign(q, 0.0);
        break;
      case BVHAR_PRIOR_GDP:  // tri.h:941-951
        group_rate.assign(in.group_rate, in.group_rate + G);
        group_rate_fac.assign(N, 1.0);
        coef_gamma_shape = in.gamma_shape; coef_gamma_rate = in.gamma_rate;
        local_lev.assign(in.local_sparsity, in.local_sparsity + N);
        contem_rate.assign(in.contem_rate, in.contem_rate + q);
        contem_gamma_shape = in.contem_gamma_shape; contem_gamma_rate = in.contem_gamma_rate;
        contem_fac.assign(in.contem_local_sparsity, in.contem_local_sparsity + q);
        break;
    }
    (void)pp;
  }

  // Record matrices and their names/order: cfg.h:639-648 (alpha, a, [c]), 865-867 (d), 967-971
  // (h, h0, sigh), 814-820 (sparse), tri.h:579-581, 671-676, 772-776, 865-868 (prior records).
  void init_records() {
    rec.cap = S->num_iter - S->num_burn;
    if (rec.cap < 0) rec.cap = 0;
    const uint32_t m = S->rec_mask;
    if (m & BVHAR_REC_COEF) {
      rec.add("alpha_record", N);
      rec.add("a_record", q);
      if (S->mean) rec.add("c_record", k);
      if (!S->sv) rec.add("d_record", k);
    }
    if (S->sv) {
      if (m & BVHAR_REC_LVOL) rec.add("h_record", n * k);
      else if (m & BVHAR_REC_LVOL_LAST) rec.add("h_last_record", k);
      if (m & BVHAR_REC_COEF) { rec.add("h0_record", k); rec.add("sigh_record", k); }
    }
    if (m & BVHAR_REC_SPARSE) {
      rec.add("alpha_sparse_record", N);
      rec.add("a_sparse_record", q);
      if (S->mean) rec.add("c_sparse_record", k);
    }
    if (m & BVHAR_REC_PRIOR) {
      switch (S->prior) {
        case BVHAR_PRIOR_SSVS: rec.add("gamma_record", N); break;
        case BVHAR_PRIOR_HORSESHOE:
          rec.add("lambda_record", N); rec.add("eta_record", G); rec.add("tau_record", 1); rec.add("kappa_record", N);
          break;
        case BVHAR_PRIOR_NG: rec.add("lambda_record", N); rec.add("eta_record", G); rec.add("tau_record", 1); break;
        case BVHAR_PRIOR_DL: rec.add("lambda_record", N); rec.add("tau_record", 1); break;
        default: break;
      }
    }
  }

  void put(const char* name, int row, const double* v, int len) {
    auto it = rec.m.find(name);
    if (it == rec.m.end()) return;
    Mat& M = it->second;
    for (int j = 0; j < len; ++j) M(row, j) = v[j];
  }

  // updateRecords: cfg.h:850-857 / 955-965 (coefficients and state), 796-804 (sparse), prior records.
  void update_records() {
    if (rec.filled >= rec.cap) return;
    const int r = rec.filled++;
    put("alpha_record", r, coef_vec.data(), N);
    if (S->mean) put("c_record", r, coef_vec.data() + N, k);
    put("a_record", r, contem_coef.data(), q);
    if (!S->sv) put("d_record", r, diag_vec.data(), k);
    else {
      auto it = rec.m.find("h_record");
      if (it != rec.m.end()) {  // lvol_draw.transpose().reshaped(): time-major, k contiguous (cfg.h:962)
        Mat& M = it->second;
        for (int t = 0; t < n; ++t)
          for (int i = 0; i < k; ++i) M(r, t * k + i) = lvol_draw(t, i);
      }
      it = rec.m.find("h_last_record");
      if (it != rec.m.end())
        for (int i = 0; i < k; ++i) it->second(r, i) = lvol_draw(n - 1, i);
      put("h0_record", r, lvol_init.data(), k);
      put("sigh_record", r, lvol_sig.data(), k);
    }
    if (rec.m.count("alpha_sparse_record")) {
      Mat& M = rec.m["alpha_sparse_record"];
      for (int j = 0; j < k; ++j)
        for (int i = 0; i < nrow; ++i) M(r, j * nrow + i) = sparse_coef(i, j);
      if (S->mean) {
        Mat& Cc = rec.m["c_sparse_record"];
        for (int j = 0; j < k; ++j) Cc(r, j) = sparse_coef(nrow, j);
      }
      put("a_sparse_record", r, sparse_contem.data(), q);
    }
    switch (S->prior) {
      case BVHAR_PRIOR_SSVS: put("gamma_record", r, coef_dummy.data(), N); break;
      case BVHAR_PRIOR_HORSESHOE:
        put("lambda_record", r, local_lev.data(), N); put("eta_record", r, group_lev.data(), G);
        put("tau_record", r, &global_lev, 1); put("kappa_record", r, shrink_fac.data(), N);
        break;
      case BVHAR_PRIOR_NG:
        put("lambda_record", r, local_lev.data(), N); put("eta_record", r, group_lev.data(), G);
        put("tau_record", r, &global_lev, 1);
        break;
      case BVHAR_PRIOR_DL:
        put("lambda_record", r, local_lev.data(), N); put("tau_record", r, &global_lev, 1);
        break;
      default: break;
    }
  }

  // ---------------------------------------------------------------- the sweep (tri.h:85-115)
  void sweep(bool record) {
    ++sweep_count;
    rng.iter = sweep_count;
    for (int s = 1; s <= 9; ++s) stage(s);
    if (record) update_records();
  }
  void stage(int s) {
    switch (s) {
      case 1: update_coef_prec(); break;
      case 2: update_penalty(); break;
      case 3: update_sv(); break;
      case 4: if (S->faithful) update_coef_faithful(); else update_coef_efficient(); break;
      case 5: update_impact_prec(); break;
      case 6: update_latent(); break;
      case 7: update_impact(); break;
      case 8: chol_lower = build_inv_lower(k, contem_coef.data()); break;  // tri.h:348
      case 9: if (S->sv) update_state_sv(); else update_state_ldlt(); break;
    }
  }

  // tri.h:254-262
  void update_penalty() {
    for (int i = 0; i < N; ++i) alpha_penalty[i] = S->own.count(S->grp_vec[i]) ? 0.0 : 1.0;
  }
  // tri.h:372 (LDLT), tri.h:409 (SV)
  void update_sv() {
    if (!S->sv) {
      for (int i = 0; i < k; ++i) { const double s = std::sqrt(diag_vec[i]); for (int t = 0; t < n; ++t) sqrt_sv(t, i) = s; }
    } else {
      for (int i = 0; i < k; ++i) for (int t = 0; t < n; ++t) sqrt_sv(t, i) = std::exp(lvol_draw(t, i) / 2);
    }
  }
  // tri.h:342
  void update_latent() {
    const Mat& X = Wd->X; const Mat& Y = Wd->Y;
    for (int j = 0; j < k; ++j) {
      double* z = latent_innov.col(j);
      const double* y = Y.col(j);
      for (int t = 0; t < n; ++t) z[t] = y[t];
      for (int a = 0; a < d; ++a) {
        const double c = coef_mat(a, j);
        if (c == 0.0) continue;
        const double* x = X.col(a);
        for (int t = 0; t < n; ++t) z[t] -= x[t] * c;
      }
    }
  }

  void prior_slice(int j, Vec& mean_j, Vec& prec_j, Vec& pen_j) const {  // tri.h:287-305
    mean_j.assign(d, 0.0); prec_j.assign(d, 0.0); pen_j.assign(d, 0.0);
    if (S->mean) {
      for (int i = 0; i < nrow; ++i) {
        mean_j[i] = prior_alpha_mean[(size_t)j * nrow + i];
        prec_j[i] = prior_alpha_prec[(size_t)j * nrow + i];
        pen_j[i] = alpha_penalty[(size_t)j * nrow + i];
      }
      mean_j[nrow] = prior_alpha_mean[N + j];
      prec_j[nrow] = prior_alpha_prec[N + j];
    } else {
      for (int i = 0; i < d; ++i) {
        mean_j[i] = prior_alpha_mean[(size_t)j * d + i];
        prec_j[i] = prior_alpha_prec[(size_t)j * d + i];
        pen_j[i] = alpha_penalty[(size_t)j * d + i];
      }
    }
  }
  // draw_mn_savs, draw.h:417-430 (0/0 -> NaN when the coefficient is exactly 0, as in the reference)
  void mn_savs(int j, const Vec& pen_j) {
    for (int i = 0; i < d; ++i) {
      const double c = coef_mat(i, j);
      const double pen = pen_j[i] / (c * c);
      const double fit = std::fabs(c) * Wd->xnorm[i];
      sparse_coef(i, j) = c * std::max(fit - pen, 0.0) / fit;
    }
  }

  // updateCoef exactly as written, tri.h:281-316 + draw_coef draw.h:372-383.
  void update_coef_faithful() {
    const Mat& X = Wd->X; const Mat& Y = Wd->Y;
    Vec mean_j, prec_j, pen_j, z(d), rhs(d), out(d);
    for (int j = 0; j < k; ++j) {
      for (int a = 0; a < d; ++a) coef_mat(a, j) = 0.0;
      const int kk = k - j;
      // design_coef = kron(L[j:, j], X) with row (i-j)*n + t divided by sqrt_sv(t, i)     (tri.h:286)
      Mat design(n * kk, d);
      for (int a = 0; a < d; ++a)
        for (int i = j; i < k; ++i)
          for (int t = 0; t < n; ++t) design((i - j) * n + t, a) = chol_lower(i, j) * X(t, a) / sqrt_sv(t, i);
      // response = (((Y - X A) L[j:,:]^T) / sqrt_sv[:, j:]).reshaped()                    (tri.h:297)
      Mat E(n, k);
      {
        Mat XA; gemm_nn(X, coef_mat, XA);
        for (size_t e = 0; e < E.size(); ++e) E.v[e] = Y.v[e] - XA.v[e];
      }
      Vec resp((size_t)n * kk);
      for (int i = j; i < k; ++i)
        for (int t = 0; t < n; ++t) {
          double s = 0;
          for (int m = 0; m < k; ++m) s += E(t, m) * chol_lower(i, m);
          resp[(size_t)(i - j) * n + t] = s / sqrt_sv(t, i);
        }
      prior_slice(j, mean_j, prec_j, pen_j);
      rng.normal_vec(ST_COEF_Z, (uint32_t)(j * d), d, z.data());  // draw.h:375-378: z first
      Mat P; gemm_tn(design, design, P);                         // x' x
      for (int a = 0; a < d; ++a) P(a, a) += prec_j[a];
      for (int a = 0; a < d; ++a) {
        double s = 0;
        const double* dc = design.col(a);
        for (size_t r = 0; r < resp.size(); ++r) s += dc[r] * resp[r];
        rhs[a] = prec_j[a] * mean_j[a] + s;
      }
      precision_draw(P, rhs.data(), z.data(), out.data());
      for (int a = 0; a < d; ++a) coef_mat(a, j) = out[a];
      refresh_coef_vec();
      mn_savs(j, pen_j);
    }
  }

  // Same conditional draws through the sufficient statistics S_i = X' D_i X, C_i = X' D_i (Y L_i'),
  // D_i = diag(1 / sqrt_sv(:, i)^2)  (DESIGN.md "updateCoef"):
  //   P_j = diag(prec_j) + sum_{i>=j} L_ij^2 S_i
  //   b_j = prec_j o mean_j + sum_{i>=j} L_ij (C_i - S_i v_i),  v_i = sum_{m != j} L_im a_m
  void update_coef_efficient() {
    const Mat& X = Wd->X; const Mat& Y = Wd->Y;
    std::vector<Mat> Sm(k);
    Mat Cm(d, k);
    if (S->sv) {
      for (int i = 0; i < k; ++i) {
        Mat XD(n, d);
        for (int a = 0; a < d; ++a)
          for (int t = 0; t < n; ++t) XD(t, a) = X(t, a) / (sqrt_sv(t, i) * sqrt_sv(t, i));
        gemm_tn(XD, X, Sm[i]);
        for (int t = 0; t < n; ++t) {
          double yl = 0;
          for (int m = 0; m <= i; ++m) yl += chol_lower(i, m) * Y(t, m);
          for (int a = 0; a < d; ++a) Cm(a, i) += XD(t, a) * yl;
        }
      }
    } else {  // homoskedastic: S_i = X'X / d_i, C_i = (X'Y) L_i' / d_i
      for (int i = 0; i < k; ++i) {
        Sm[i] = Wd->XtX;
        const double inv = 1.0 / (sqrt_sv(0, i) * sqrt_sv(0, i));
        for (auto& x : Sm[i].v) x *= inv;
        for (int a = 0; a < d; ++a) {
          double s = 0;
          for (int m = 0; m <= i; ++m) s += Wd->XtY(a, m) * chol_lower(i, m);
          Cm(a, i) = s * inv;
        }
      }
    }
    Mat V(d, k);  // v_i = A L_i'
    for (int i = 0; i < k; ++i)
      for (int m = 0; m <= i; ++m) {
        const double l = chol_lower(i, m);
        for (int a = 0; a < d; ++a) V(a, i) += l * coef_mat(a, m);
      }
    Vec mean_j, prec_j, pen_j, z(d), rhs(d), out(d), tmp(d);
    for (int j = 0; j < k; ++j) {
      for (int i = j; i < k; ++i) {
        const double l = chol_lower(i, j);
        for (int a = 0; a < d; ++a) V(a, i) -= l * coef_mat(a, j);
      }
      for (int a = 0; a < d; ++a) coef_mat(a, j) = 0.0;
      prior_slice(j, mean_j, prec_j, pen_j);
      rng.normal_vec(ST_COEF_Z, (uint32_t)(j * d), d, z.data());
      Mat P(d, d);
      for (int a = 0; a < d; ++a) { P(a, a) = prec_j[a]; rhs[a] = prec_j[a] * mean_j[a]; }
      for (int i = j; i < k; ++i) {
        const double l = chol_lower(i, j), l2 = l * l;
        for (size_t e = 0; e < P.size(); ++e) P.v[e] += l2 * Sm[i].v[e];
        for (int a = 0; a < d; ++a) tmp[a] = 0.0;
        for (int b = 0; b < d; ++b) {
          const double vb = V(b, i);
          const double* sc = Sm[i].col(b);
          for (int a = 0; a < d; ++a) tmp[a] += sc[a] * vb;
        }
        for (int a = 0; a < d; ++a) rhs[a] += l * (Cm(a, i) - tmp[a]);
      }
      precision_draw(P, rhs.data(), z.data(), out.data());
      for (int a = 0; a < d; ++a) coef_mat(a, j) = out[a];
      for (int i = j; i < k; ++i) {
        const double l = chol_lower(i, j);
        for (int a = 0; a < d; ++a) V(a, i) += l * out[a];
      }
      refresh_coef_vec();
      mn_savs(j, pen_j);
    }
  }

  // updateImpact tri.h:322-336 + draw_coef + draw_savs draw.h:398-415
  void update_impact() {
    Vec z, rhs, out;
    for (int j = 1; j < k; ++j) {
      const int id = j * (j - 1) / 2;
      Mat design(n, j);
      Vec resp(n);
      for (int t = 0; t < n; ++t) resp[t] = latent_innov(t, j) / sqrt_sv(t, j);
      for (int b = 0; b < j; ++b)
        for (int t = 0; t < n; ++t) design(t, b) = latent_innov(t, b) / sqrt_sv(t, j);
      z.assign(j, 0.0); rhs.assign(j, 0.0); out.assign(j, 0.0);
      rng.normal_vec(ST_IMPACT_Z, (uint32_t)id, j, z.data());
      Mat P; gemm_tn(design, design, P);
      for (int b = 0; b < j; ++b) {
        P(b, b) += prior_chol_prec[id + b];
        double s = 0;
        for (int t = 0; t < n; ++t) s += design(t, b) * resp[t];
        rhs[b] = prior_chol_prec[id + b] * prior_chol_mean[id + b] + s;
      }
      precision_draw(P, rhs.data(), z.data(), out.data());
      for (int b = 0; b < j; ++b) {
        contem_coef[id + b] = out[b];
        // draw_savs: penalty 1/a^2, norms of the *unscaled* latent columns
        double nrm = 0;
        for (int t = 0; t < n; ++t) nrm += latent_innov(t, b) * latent_innov(t, b);
        const double fit = std::fabs(out[b]) * nrm;
        sparse_contem[id + b] = out[b] * std::max(fit - 1.0 / (out[b] * out[b]), 0.0) / fit;
      }
    }
  }

  Mat ortho_latent() const {  // latent_innov * chol_lower^T (tri.h:371, 401)
    Mat U(n, k);
    for (int i = 0; i < k; ++i)
      for (int m = 0; m <= i; ++m) {
        const double l = chol_lower(i, m);
        if (l == 0.0) continue;
        for (int t = 0; t < n; ++t) U(t, i) += l * latent_innov(t, m);
      }
    return U;
  }

  // McmcReg::updateState tri.h:371 / reg_ldlt_diag draw.h:1244-1259 (num_design / 2 is integer division)
  void update_state_ldlt() {
    Mat U = ortho_latent();
    for (int i = 0; i < k; ++i) {
      double ss = 0;
      for (int t = 0; t < n; ++t) ss += U(t, i) * U(t, i);
      diag_vec[i] = 1.0 / rng.gamma(ST_LDLT_DIAG, (uint32_t)i, S->sig_shp[i] + n / 2, 1.0 / (S->sig_scl[i] + ss / 2));
    }
  }

  // Mixture indicator of one time point, draw.h:461-468 (inverse-CDF on the normalised weights).
  static int ksc_indicator(double ystar, double h, double u) {
    double pdf[7], tot = 0;
    for (int c = 0; c < 7; ++c) {
      const double m = KSC_M[c] - 1.2704, sd = std::sqrt(KSC_V[c]);
      const double e = (ystar - h - m) / sd;
      pdf[c] = std::exp(-e * e / 2) * KSC_P[c] / (sd * std::sqrt(2 * M_PI));
      tot += pdf[c];
    }
    double cum = 0;
    int cnt = 0;  // number of components whose cumulative weight exceeds u
    for (int c = 0; c < 7; ++c) { cum += pdf[c] / tot; if (u < cum) ++cnt; }
    int s = 7 - cnt;  // draw.h:468
    return s > 6 ? 6 : s;
  }

  // varsv_ht exactly as written: dense n x n, draw.h:440-488.
  void varsv_ht_faithful(int i, const double* ystar, const double* u, const double* z) {
    const double sig = lvol_sig[i], h0 = lvol_init[i];
    std::vector<int> s(n);
    for (int t = 0; t < n; ++t) s[t] = ksc_indicator(ystar[t], lvol_draw(t, i), u[t]);
    Mat diff(n, n);
    for (int t = 0; t < n; ++t) diff(t, t) = 1.0;
    for (int t = 0; t + 1 < n; ++t) diff(t + 1, t) = -1.0;
    Mat HtH; gemm_tn(diff, diff, HtH);
    Mat post(n, n);
    for (size_t e = 0; e < post.size(); ++e) post.v[e] = HtH.v[e] / sig;
    Vec rhs(n, 0.0);
    for (int t = 0; t < n; ++t) {
      post(t, t) += 1.0 / KSC_V[s[t]];
      double hh = 0;  // (HtH * init_sv * ones / sv_sig)[t]
      for (int c = 0; c < n; ++c) hh += HtH(t, c) * h0;
      rhs[t] = hh / sig + (ystar[t] - (KSC_M[s[t]] - 1.2704)) / KSC_V[s[t]];
    }
    Vec out(n);
    precision_draw(post, rhs.data(), z, out.data());
    for (int t = 0; t < n; ++t) lvol_draw(t, i) = out[t];
  }
  // Same draw through the tridiagonal structure of H'H/sig + diag(1/v_s).
  void varsv_ht_efficient(int i, const double* ystar, const double* u, const double* z) {
    const double isig = 1.0 / lvol_sig[i], h0 = lvol_init[i];
    Vec l(n), e(n), y(n);
    for (int t = 0; t < n; ++t) {
      const int s = ksc_indicator(ystar[t], lvol_draw(t, i), u[t]);
      const double iv = 1.0 / KSC_V[s];
      const double dg = (t < n - 1 ? 2.0 : 1.0) * isig + iv;
      const double b = (t == 0 ? h0 * isig : 0.0) + (ystar[t] - (KSC_M[s] - 1.2704)) * iv;
      if (t == 0) { e[0] = 0; l[0] = std::sqrt(dg); y[0] = b / l[0]; }
      else { e[t] = -isig / l[t - 1]; l[t] = std::sqrt(dg - e[t] * e[t]); y[t] = (b - e[t] * y[t - 1]) / l[t]; }
    }
    double next = 0;
    for (int t = n - 1; t >= 0; --t) {  // L^T h = y + z
      double v = y[t] + z[t];
      if (t < n - 1) v -= e[t + 1] * next;
      next = v / l[t];
      lvol_draw(t, i) = next;
    }
  }

  // McmcSv::updateState tri.h:400-408; varsv_sigh draw.h:498-512; varsv_h0 draw.h:523-537
  void update_state_sv() {
    Mat U = ortho_latent();
    for (auto& x : U.v) x = std::log(x * x + .0001);
    // Series are independent given the orthogonalised residuals; with the addressed (Philox) stream
    // they may run concurrently when fewer units than threads are being swept (inner_threads > 1).
#pragma omp parallel for num_threads(inner_threads) schedule(dynamic, 1) if (inner_threads > 1 && rng.philox)
    for (int i = 0; i < k; ++i) {
      Vec u(n), z(n), ys(n);
      if (rng.philox) {
        // element (series i, time t) = i * 2*ceil(n/2) + t: the two lanes of a Philox pair are
        // consecutive time points of the same series (DESIGN.md "Random streams")
        const uint32_t base = (uint32_t)i * 2u * (uint32_t)((n + 1) / 2);
        rng.uniform_vec(ST_SV_MIX_U, base, n, u.data());
        rng.normal_vec(ST_SV_H_Z, base, n, z.data());
      } else {  // reference order: n uniforms then n normals (draw.h:458-460, 479-481)
        rng.uniform_vec(ST_SV_MIX_U, 0, n, u.data());
        rng.normal_vec(ST_SV_H_Z, 0, n, z.data());
      }
      for (int t = 0; t < n; ++t) ys[t] = U(t, i);
      if (S->faithful) varsv_ht_faithful(i, ys.data(), u.data(), z.data());
      else varsv_ht_efficient(i, ys.data(), u.data(), z.data());
    }
    // varsv_sigh: the sum of squared increments runs over *all* series (draw.h:508); n / 2 integer.
    double ss = 0;
    for (int i = 0; i < k; ++i)
      for (int t = 0; t < n; ++t) {
        const double prev = t == 0 ? lvol_init[i] : lvol_draw(t - 1, i);
        const double df = lvol_draw(t, i) - prev;
        ss += df * df;
      }
    for (int i = 0; i < k; ++i)
      lvol_sig[i] = 1.0 / rng.gamma(ST_SV_SIGH, (uint32_t)i, S->sig_shp[i] + n / 2, 1.0 / (S->sig_scl[i] + ss / 2));
    // varsv_h0
    Mat K = S->sv_prec;
    Vec rhs(k, 0.0), zz(k), out(k);
    for (int i = 0; i < k; ++i) {
      K(i, i) += 1.0 / lvol_sig[i];
      double s = 0;
      for (int m = 0; m < k; ++m) s += S->sv_prec(i, m) * S->sv_mean[m];
      rhs[i] = s + lvol_draw(0, i) / lvol_sig[i];
    }
    rng.normal_vec(ST_SV_H0_Z, 0, k, zz.data());
    precision_draw(K, rhs.data(), zz.data(), out.data());
    lvol_init = out;
  }

  // ---------------------------------------------------------------- priors
  static void griddy_weights(Vec& logw, Vec& w) {  // exp(logw - max), NaN -> 0
    double mx = -std::numeric_limits<double>::infinity();
    for (double v : logw) if (v == v && v > mx) mx = v;
    w.resize(logw.size());
    for (size_t i = 0; i < logw.size(); ++i) {
      const double e = std::exp(logw[i] - mx);
      w[i] = (e == e) ? e : 0.0;
    }
  }
  static double interior_grid(int grid, int i) {  // LinSpaced(grid + 2, 0, 1).segment(1, grid)[i]
    return (double)(i + 1) * (1.0 / (double)(grid + 1));  // Eigen: low + i * step
  }

  // horseshoe_global_sparsity draw.h:666-677 ((dim + 1) / 2 is integer division)
  double hs_global(uint32_t st, uint32_t idx, double latent, const Vec& local, const Vec& coef) {
    double s = 0;
    for (size_t i = 0; i < coef.size(); ++i) s += coef[i] * coef[i] / (2 * local[i] * local[i]);
    return std::sqrt(1.0 / rng.gamma(st, idx, (double)(((int)coef.size() + 1) / 2), 1.0 / (1.0 / latent + s)));
  }
  // ng_shape_jump draw.h:1098-1119 (lgammafn = R's lgamma; std::lgamma in the Python build, defs.h:66-68)
  double ng_shape_jump(uint32_t st_z, uint32_t st_u, uint32_t idx, double gam, const Vec& local, double glob, double sd) {
    const int nc = (int)local.size();
    const double cand = std::exp(std::log(gam) + rng.normal(st_z, idx) * sd);
    double slog = 0, ssq = 0;
    for (double v : local) { slog += std::log(v); ssq += v * v; }
    double lr = std::log(cand) - std::log(gam) + nc * (std::lgamma(gam) - std::lgamma(cand));
    lr += nc * cand * (std::log(cand) - 2 * std::log(glob));
    lr -= nc * gam * (std::log(gam) - 2 * std::log(glob));
    lr += (cand - gam) * slog;
    lr += (gam - cand) * ssq / (glob * glob);
    if (std::log(rng.uniform(st_u, idx)) < std::min(lr, 0.0)) return cand;
    return gam;
  }
  // dl_dir_griddy draw.h:880-890 (1 / size() is integer division -> 0 for size > 1)
  double dl_dir_griddy(uint32_t st, int grid, const Vec& local, double glob) {
    const int nc = (int)local.size();
    const double lo0 = (double)(1 / nc);
    const double lo = lo0 < .5 ? lo0 : .5, hi = lo0 < .5 ? .5 : lo0;
    double slog = 0;
    for (double v : local) slog += std::log(v);
    Vec logw(grid), w, gridv(grid);
    for (int i = 0; i < grid; ++i) {
      const double step = grid > 1 ? (hi - lo) / (grid - 1) : 0.0;
      const double c = i == grid - 1 ? hi : lo + i * step;  // Eigen LinSpaced: low + i*step, last = high
      gridv[i] = c;
      logw[i] = c * (nc * (std::log(glob) - std::log(2.0)) + slog) - nc * std::lgamma(c);  // draw.h:870
    }
    griddy_weights(logw, w);
    return gridv[rng.categorical(st, 0, w.data(), grid)];
  }

  void update_coef_prec() {
    const bvhar_prior_params& pp = S->pp;
    const uint32_t P0 = ST_PRIOR_COEF;
    Vec coef(coef_vec.begin(), coef_vec.begin() + N);
    const std::vector<int>& gv = S->grp_vec;
    switch (S->prior) {
      case BVHAR_PRIOR_MINNESOTA: break;  // tri.h:455
      case BVHAR_PRIOR_HIERMINN: {        // tri.h:513-524
        // minnesota_lambda draw.h:956-965 (coef.size() / 2 integer)
        double chi = 0;
        for (int i = 0; i < N; ++i) {
          prior_alpha_prec[i] *= own_lambda;
          const double df = coef[i] - prior_alpha_mean[i];
          chi += df * df * prior_alpha_prec[i];
        }
        own_lambda = rng.gig(P0 + 0, 0, pp.hmn_shape - N / 2, 2 * pp.hmn_rate, chi);
        cut_param(own_lambda);
        for (int i = 0; i < N; ++i) prior_alpha_prec[i] /= own_lambda;
        // minnesota_nu_griddy draw.h:981-998 with an empty cross_id set (draw.h:75): the weights are
        // uniform, nu is drawn and touches no coefficient.
        Vec w(pp.hmn_grid_size, 1.0);
        cross_lambda = interior_grid(pp.hmn_grid_size, rng.categorical(P0 + 1, 0, w.data(), pp.hmn_grid_size));
        break;
      }
      case BVHAR_PRIOR_SSVS: {  // tri.h:599-616
        for (int i = 0; i < N; ++i)  // ssvs_local_slab draw.h:336-345
          coef_slab[i] = std::sqrt(1.0 / rng.gamma(P0 + 0, (uint32_t)i, pp.ssvs_coef_slab_shape + .5,
                         1.0 / (pp.ssvs_coef_slab_scl + coef[i] * coef[i] / (coef_dummy[i] + (1 - coef_dummy[i]) * spike_scl))));
        for (int i = 0; i < N; ++i)
          for (int g = 0; g < G; ++g) if (gv[i] == S->grp_id[g]) slab_weight[i] = coef_weight[g];
        {  // ssvs_scl_griddy draw.h:347-361
          const int grid = pp.ssvs_coef_grid;
          double ss = 0;
          for (int i = 0; i < N; ++i) { const double r = coef[i] / coef_slab[i]; ss += r * r; }
          Vec logw(grid), w;
          for (int i = 0; i < grid; ++i) { const double c = interior_grid(grid, i); logw[i] = -ss / (2 * c * c) - N * std::log(c); }
          griddy_weights(logw, w);
          spike_scl = interior_grid(grid, rng.categorical(P0 + 1, 0, w.data(), grid));
        }
        for (int i = 0; i < N; ++i) {  // ssvs_dummy draw.h:269-281
          const double sn = coef_slab[i], sdn = spike_scl * coef_slab[i];
          double e1 = -coef[i] * coef[i] / (2 * sn * sn), e2 = -coef[i] * coef[i] / (2 * sdn * sdn);
          const double mx = std::max(e1, e2);
          e1 = slab_weight[i] * std::exp(e1 - mx) / sn;
          e2 = (1 - slab_weight[i]) * std::exp(e2 - mx) / sdn;
          coef_dummy[i] = rng.bernoulli(P0 + 2, (uint32_t)i, e1 / (e1 + e2));
        }
        for (int g = 0; g < G; ++g) {  // ssvs_mn_weight draw.h:308-325
          double sm = 0; int cnt = 0;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) { sm += coef_dummy[i]; ++cnt; }
          coef_weight[g] = rng.beta(P0 + 3, (uint32_t)g, S->ssvs_s1[g] + sm, S->ssvs_s2[g] + cnt - sm);
        }
        for (int i = 0; i < N; ++i)  // tri.h:615 (first power, not squared)
          prior_alpha_prec[i] = 1.0 / (spike_scl * (1 - coef_dummy[i]) * coef_slab[i] + coef_dummy[i] * coef_slab[i]);
        break;
      }
      case BVHAR_PRIOR_HORSESHOE: {  // tri.h:693-711
        for (int g = 0; g < G; ++g)  // horseshoe_latent draw.h:740-745
          latent_group[g] = 1.0 / rng.gamma(P0 + 0, (uint32_t)g, 1.0, 1.0 / (1 + 1 / (group_lev[g] * group_lev[g])));
        for (int g = 0; g < G; ++g) {  // horseshoe_mn_sparsity draw.h:713-733
          Vec mc, ml;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) { mc.push_back(coef[i]); ml.push_back(global_lev * local_lev[i]); }
          group_lev[g] = hs_global(P0 + 1, (uint32_t)g, latent_group[g], ml, mc);
        }
        for (int i = 0; i < N; ++i)
          for (int g = 0; g < G; ++g) if (gv[i] == S->grp_id[g]) coef_var[i] = group_lev[g];
        for (int i = 0; i < N; ++i)
          latent_local[i] = 1.0 / rng.gamma(P0 + 2, (uint32_t)i, 1.0, 1.0 / (1 + 1 / (local_lev[i] * local_lev[i])));
        if (S->ggl) {
          latent_global = 1.0 / rng.gamma(P0 + 3, 0, 1.0, 1.0 / (1 + 1 / (global_lev * global_lev)));
          Vec lh(N);
          for (int i = 0; i < N; ++i) lh[i] = coef_var[i] * local_lev[i];
          global_lev = hs_global(P0 + 4, 0, latent_global, lh, coef);
        }
        for (int i = 0; i < N; ++i) {  // horseshoe_local_sparsity draw.h:649-656, prior_var = global^2
          const double scl = 1.0 / (1 / latent_local[i] + coef[i] * coef[i] / (2 * global_lev * global_lev * coef_var[i] * coef_var[i]));
          local_lev[i] = std::sqrt(1.0 / rng.gamma(P0 + 5, (uint32_t)i, 1.0, scl));
        }
        for (int i = 0; i < N; ++i) {
          const double sdv = global_lev * coef_var[i] * local_lev[i];
          prior_alpha_prec[i] = 1.0 / (sdv * sdv);
          shrink_fac[i] = 1.0 / (1 + prior_alpha_prec[i]);
        }
        break;
      }
      case BVHAR_PRIOR_NG: {  // tri.h:793-812
        for (int g = 0; g < G; ++g) {  // ng_mn_shape_jump draw.h:1121-1139
          Vec ml;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) ml.push_back(local_lev[i]);
          local_shape[g] = ng_shape_jump(P0 + 0, P0 + 1, (uint32_t)g, local_shape[g], ml, global_lev * group_lev[g], pp.ng_shape_sd);
        }
        for (int g = 0; g < G; ++g) {  // ng_mn_sparsity draw.h:1076-1095; scl is global_scale (tri.h:762)
          double ssq = 0; int cnt = 0;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) { const double v = local_lev[i] / global_lev; ssq += v * v; ++cnt; }
          double t = std::sqrt(1.0 / rng.gamma(P0 + 2, (uint32_t)g, pp.ng_group_shape + cnt * local_shape[g],
                                               1.0 / (local_shape[g] * ssq + pp.ng_global_scale)));
          cut_param(t);
          group_lev[g] = t;
        }
        for (int i = 0; i < N; ++i)
          for (int g = 0; g < G; ++g) if (gv[i] == S->grp_id[g]) { coef_var[i] = group_lev[g]; local_shape_fac[i] = local_shape[g]; }
        if (S->ggl) {  // ng_global_sparsity (vector overload) draw.h:1056-1070
          double sg = 0, ssq = 0;
          for (int i = 0; i < N; ++i) { sg += local_shape_fac[i]; const double v = local_lev[i] / coef_var[i]; ssq += local_shape_fac[i] * v * v; }
          double t = std::sqrt(1.0 / rng.gamma(P0 + 3, 0, pp.ng_global_shape + sg, 1.0 / (ssq + pp.ng_global_scale)));
          cut_param(t);
          global_lev = t;
        }
        for (int i = 0; i < N; ++i) {  // ng_local_sparsity draw.h:1020-1031
          const double gl = global_lev * coef_var[i];
          double t = std::sqrt(rng.gig(P0 + 4, (uint32_t)i, local_shape_fac[i] - .5, 2 * local_shape_fac[i] / (gl * gl), coef[i] * coef[i]));
          cut_param(t);
          local_lev[i] = t;
        }
        for (int i = 0; i < N; ++i) prior_alpha_prec[i] = 1.0 / (local_lev[i] * local_lev[i]);
        break;
      }
      case BVHAR_PRIOR_DL: {  // tri.h:885-901
        for (int g = 0; g < G; ++g) {  // dl_mn_sparsity (first overload) draw.h:808-834
          double sm = 0; int cnt = 0;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) { sm += std::fabs(coef[i]) / (global_lev * local_lev[i]); ++cnt; }
          double t = 1.0 / rng.gamma(P0 + 0, (uint32_t)g, pp.dl_shape + cnt, 1.0 / (pp.dl_scale + sm));
          cut_param(t);
          group_lev[g] = t;
        }
        for (int i = 0; i < N; ++i)
          for (int g = 0; g < G; ++g) if (gv[i] == S->grp_id[g]) coef_var[i] = group_lev[g];
        dir_concen = dl_dir_griddy(P0 + 1, pp.dl_grid_size, local_lev, global_lev);
        {  // dl_local_sparsity draw.h:778-785
          double sm = 0;
          for (int i = 0; i < N; ++i) {
            double t = rng.gig(P0 + 2, (uint32_t)i, dir_concen - 1, 1.0, 2 * std::fabs(coef[i] / coef_var[i]));
            cut_param(t);
            local_lev[i] = t; sm += t;
          }
          for (int i = 0; i < N; ++i) local_lev[i] /= sm;
        }
        if (S->ggl) {  // dl_global_sparsity draw.h:793-799
          double sm = 0;
          for (int i = 0; i < N; ++i) sm += std::fabs(coef[i]) / (local_lev[i] * coef_var[i]);
          double t = rng.gig(P0 + 3, 0, N * (dir_concen - 1), 1.0, 2 * sm);
          cut_param(t);
          global_lev = t;
        }
        for (int i = 0; i < N; ++i) {  // dl_latent draw.h:758-770
          const double sc = global_lev * local_lev[i] * coef_var[i];
          double t = 1.0 / rng.invgauss(P0 + 4, (uint32_t)i, sc / std::fabs(coef[i]), 1.0);
          cut_param(t);
          latent_local[i] = t;
          prior_alpha_prec[i] = 1.0 / (sc * sc * t);
        }
        break;
      }
      case BVHAR_PRIOR_GDP: {  // tri.h:974-986
        {  // gdp_shape_griddy draw.h:1213-1223 (signed coefficient inside log1p, draw.h:1199)
          const int grid = pp.gdp_grid_shape;
          double sl = 0;
          for (int i = 0; i < N; ++i) sl += std::log1p(coef[i] / coef_gamma_rate);
          Vec logw(grid), w;
          for (int i = 0; i < grid; ++i) { const double c = interior_grid(grid, i); logw[i] = N * (std::log(1 - c) - std::log(c)) - sl / c; }
          griddy_weights(logw, w);
          const double c = interior_grid(grid, rng.categorical(P0 + 0, 0, w.data(), grid));
          coef_gamma_shape = (1 - c) / c;
        }
        {  // gdp_rate_griddy draw.h:1225-1235
          const int grid = pp.gdp_grid_rate;
          Vec logw(grid), w;
          for (int g = 0; g < grid; ++g) {
            const double c = interior_grid(grid, g);
            double sl = 0;
            for (int i = 0; i < N; ++i) sl += std::log1p(coef[i] * c / (1 - c));
            logw[g] = N * (std::log(c) - std::log(1 - c)) - (coef_gamma_shape + 1) * sl;
          }
          griddy_weights(logw, w);
          const double c = interior_grid(grid, rng.categorical(P0 + 1, 0, w.data(), grid));
          coef_gamma_rate = (1 - c) / c;
        }
        for (int g = 0; g < G; ++g) {  // gdp_exp_rate (group overload) draw.h:1173-1190
          double l1 = 0; int cnt = 0;
          for (int i = 0; i < N; ++i) if (gv[i] == S->grp_id[g]) { l1 += std::fabs(coef[i]); ++cnt; }
          group_rate[g] = rng.gamma(P0 + 2, (uint32_t)g, coef_gamma_shape + cnt, 1.0 / (coef_gamma_rate + l1));
        }
        for (int i = 0; i < N; ++i)
          for (int g = 0; g < G; ++g) if (gv[i] == S->grp_id[g]) group_rate_fac[i] = group_rate[g];
        for (int i = 0; i < N; ++i) {  // gdp_local_sparsity draw.h:1147-1153
          local_lev[i] = 1.0 / rng.invgauss(P0 + 3, (uint32_t)i, std::fabs(group_rate_fac[i] / coef[i]), group_rate_fac[i] * group_rate_fac[i]);
          prior_alpha_prec[i] = 1.0 / local_lev[i];
        }
        break;
      }
    }
  }

  void update_impact_prec() {
    const bvhar_prior_params& pp = S->pp;
    const uint32_t P1 = ST_PRIOR_CONTEM;
    const Vec& a = contem_coef;
    switch (S->prior) {
      case BVHAR_PRIOR_MINNESOTA: break;
      case BVHAR_PRIOR_HIERMINN: {  // tri.h:525-531
        double chi = 0;
        for (int i = 0; i < q; ++i) {
          prior_chol_prec[i] *= contem_lambda;
          const double df = a[i] - prior_chol_mean[i];
          chi += df * df * prior_chol_prec[i];
        }
        contem_lambda = rng.gig(P1 + 0, 0, pp.hmn_shape - q / 2, 2 * pp.hmn_rate, chi);
        cut_param(contem_lambda);
        for (int i = 0; i < q; ++i) prior_chol_prec[i] /= contem_lambda;
        break;
      }
      case BVHAR_PRIOR_SSVS: {  // tri.h:617-623
        for (int i = 0; i < q; ++i)
          contem_slab[i] = std::sqrt(1.0 / rng.gamma(P1 + 0, (uint32_t)i, pp.ssvs_chol_slab_shape + .5,
                           1.0 / (pp.ssvs_chol_slab_scl + a[i] * a[i] / (contem_dummy[i] + (1 - contem_dummy[i]) * contem_spike_scl))));
        {
          const int grid = pp.ssvs_chol_grid;
          double ss = 0;
          for (int i = 0; i < q; ++i) { const double r = a[i] / contem_slab[i]; ss += r * r; }
          Vec logw(grid), w;
          for (int i = 0; i < grid; ++i) { const double c = interior_grid(grid, i); logw[i] = -ss / (2 * c * c) - q * std::log(c); }
          griddy_weights(logw, w);
          contem_spike_scl = interior_grid(grid, rng.categorical(P1 + 1, 0, w.data(), grid));
        }
        for (int i = 0; i < q; ++i) {
          const double sn = contem_slab[i], sdn = contem_spike_scl * contem_slab[i];
          double e1 = -a[i] * a[i] / (2 * sn * sn), e2 = -a[i] * a[i] / (2 * sdn * sdn);
          const double mx = std::max(e1, e2);
          e1 = contem_weight[i] * std::exp(e1 - mx) / sn;
          e2 = (1 - contem_weight[i]) * std::exp(e2 - mx) / sdn;
          contem_dummy[i] = rng.bernoulli(P1 + 2, (uint32_t)i, e1 / (e1 + e2));
        }
        {  // ssvs_weight draw.h:290-297: every element gets its own Beta draw with common parameters
          double sm = 0;
          for (int i = 0; i < q; ++i) sm += contem_dummy[i];
          for (int i = 0; i < q; ++i) contem_weight[i] = rng.beta(P1 + 3, (uint32_t)i, pp.ssvs_chol_s1 + sm, pp.ssvs_chol_s2 + q - sm);
        }
        for (int i = 0; i < q; ++i) {  // tri.h:622 + build_ssvs_sd draw.h:112-116
          const double sdv = (1 - contem_dummy[i]) * contem_spike_scl * contem_slab[i] + contem_dummy[i] * contem_slab[i];
          prior_chol_prec[i] = 1.0 / (sdv * sdv);
        }
        break;
      }
      case BVHAR_PRIOR_HORSESHOE: {  // tri.h:712-720
        for (int i = 0; i < q; ++i)
          latent_contem_local[i] = 1.0 / rng.gamma(P1 + 0, (uint32_t)i, 1.0, 1.0 / (1 + 1 / (contem_local_lev[i] * contem_local_lev[i])));
        latent_contem_global[0] = 1.0 / rng.gamma(P1 + 1, 0, 1.0, 1.0 / (1 + 1 / (contem_global_lev[0] * contem_global_lev[0])));
        for (int i = 0; i < q; ++i) contem_var[i] = contem_global_lev[0];
        for (int i = 0; i < q; ++i) {
          const double scl = 1.0 / (1 / latent_contem_local[i] + a[i] * a[i] / (2 * 1.0 * contem_var[i] * contem_var[i]));
          contem_local_lev[i] = std::sqrt(1.0 / rng.gamma(P1 + 2, (uint32_t)i, 1.0, scl));
        }
        // tri.h:717 passes latent_contem_local where the local scales are expected
        contem_global_lev[0] = hs_global(P1 + 3, 0, latent_contem_global[0], latent_contem_local, a);
        for (int i = 0; i < q; ++i) {  // contem_var still holds the previous global (tri.h:715, 719)
          const double sdv = contem_var[i] * contem_local_lev[i];
          prior_chol_prec[i] = 1.0 / (sdv * sdv);
        }
        break;
      }
      case BVHAR_PRIOR_NG: {  // tri.h:813-818
        contem_shape = ng_shape_jump(P1 + 0, P1 + 1, 0, contem_shape, contem_fac, contem_global_lev[0], pp.ng_shape_sd);
        {  // ng_global_sparsity (scalar overload) draw.h:1040-1054
          double ssq = 0;
          for (int i = 0; i < q; ++i) ssq += contem_fac[i] * contem_fac[i];
          double t = std::sqrt(1.0 / rng.gamma(P1 + 2, 0, pp.ng_contem_global_shape + q * contem_shape,
                                               1.0 / (contem_shape * ssq + pp.ng_contem_global_scale)));
          cut_param(t);
          contem_global_lev[0] = t;
        }
        for (int i = 0; i < q; ++i) {
          const double gl = contem_global_lev[0];
          double t = std::sqrt(rng.gig(P1 + 3, (uint32_t)i, contem_shape - .5, 2 * contem_shape / (gl * gl), a[i] * a[i]));
          cut_param(t);
          contem_fac[i] = t;
          prior_chol_prec[i] = 1.0 / (t * t);
        }
        break;
      }
      case BVHAR_PRIOR_DL: {  // tri.h:902-908
        contem_dir_concen = dl_dir_griddy(P1 + 0, pp.dl_grid_size, contem_local_lev, contem_global_lev[0]);
        {
          double sm = 0;
          for (int i = 0; i < q; ++i) {
            double t = rng.gig(P1 + 1, (uint32_t)i, contem_dir_concen - 1, 1.0, 2 * std::fabs(a[i]));
            cut_param(t);
            contem_local_lev[i] = t; sm += t;
          }
          for (int i = 0; i < q; ++i) contem_local_lev[i] /= sm;
        }
        {
          double sm = 0;
          for (int i = 0; i < q; ++i) sm += std::fabs(a[i]) / contem_local_lev[i];
          double t = rng.gig(P1 + 2, 0, q * (contem_dir_concen - 1), 1.0, 2 * sm);
          cut_param(t);
          contem_global_lev[0] = t;
        }
        for (int i = 0; i < q; ++i) {  // tri.h:906: the latent uses local only (no global)
          double t = 1.0 / rng.invgauss(P1 + 3, (uint32_t)i, contem_local_lev[i] / std::fabs(a[i]), 1.0);
          cut_param(t);
          latent_contem_local[i] = t;
          const double sc = contem_global_lev[0] * contem_local_lev[i];
          prior_chol_prec[i] = 1.0 / (sc * sc * t);
        }
        break;
      }
      case BVHAR_PRIOR_GDP: {  // tri.h:987-993
        {
          const int grid = pp.gdp_grid_shape;
          double sl = 0;
          for (int i = 0; i < q; ++i) sl += std::log1p(a[i] / contem_gamma_rate);
          Vec logw(grid), w;
          for (int i = 0; i < grid; ++i) { const double c = interior_grid(grid, i); logw[i] = q * (std::log(1 - c) - std::log(c)) - sl / c; }
          griddy_weights(logw, w);
          const double c = interior_grid(grid, rng.categorical(P1 + 0, 0, w.data(), grid));
          contem_gamma_shape = (1 - c) / c;
        }
        {
          const int grid = pp.gdp_grid_rate;
          Vec logw(grid), w;
          for (int g = 0; g < grid; ++g) {
            const double c = interior_grid(grid, g);
            double sl = 0;
            for (int i = 0; i < q; ++i) sl += std::log1p(a[i] * c / (1 - c));
            logw[g] = q * (std::log(c) - std::log(1 - c)) - (contem_gamma_shape + 1) * sl;
          }
          griddy_weights(logw, w);
          const double c = interior_grid(grid, rng.categorical(P1 + 1, 0, w.data(), grid));
          contem_gamma_rate = (1 - c) / c;
        }
        for (int i = 0; i < q; ++i)  // gdp_exp_rate (element overload) draw.h:1160-1165
          contem_rate[i] = rng.gamma(P1 + 2, (uint32_t)i, contem_gamma_shape + 1, 1.0 / (contem_gamma_rate + std::fabs(a[i])));
        for (int i = 0; i < q; ++i) {
          contem_fac[i] = 1.0 / rng.invgauss(P1 + 3, (uint32_t)i, std::fabs(contem_rate[i] / a[i]), contem_rate[i] * contem_rate[i]);
          prior_chol_prec[i] = 1.0 / contem_fac[i];
        }
        break;
      }
    }
  }

  // ---------------------------------------------------------------- state access by member name
  struct Ref { double* p; size_t len; };
  Ref state(const std::string& nm) {
#define V_(name) if (nm == #name) return {name.data(), name.size()};
#define S_(name) if (nm == #name) return {&name, 1};
    V_(coef_mat) V_(contem_coef) V_(chol_lower) V_(latent_innov) V_(sqrt_sv) V_(prior_alpha_mean) V_(prior_alpha_prec)
    V_(alpha_penalty) V_(prior_chol_prec) V_(sparse_coef) V_(sparse_contem) V_(coef_vec) V_(diag_vec) V_(lvol_draw)
    V_(lvol_init) V_(lvol_sig) V_(coef_dummy) V_(coef_weight) V_(contem_dummy) V_(contem_weight) V_(coef_slab)
    V_(contem_slab) V_(local_lev) V_(group_lev) V_(shrink_fac) V_(latent_local) V_(latent_group) V_(coef_var)
    V_(contem_local_lev) V_(contem_global_lev) V_(latent_contem_local) V_(latent_contem_global) V_(local_shape)
    V_(contem_fac) V_(group_rate) V_(contem_rate)
    S_(own_lambda) S_(cross_lambda) S_(contem_lambda) S_(spike_scl) S_(contem_spike_scl) S_(global_lev) S_(latent_global)
    S_(contem_shape) S_(dir_concen) S_(contem_dir_concen) S_(coef_gamma_shape) S_(coef_gamma_rate)
    S_(contem_gamma_shape) S_(contem_gamma_rate)
#undef V_
#undef S_
    return {nullptr, 0};
  }
};

struct Fit {
  Shared S;
  std::vector<Window> wins;
  std::vector<std::unique_ptr<Chain>> chains;  // unit = w * C + c
  int nthreads = 0;
};

static Fit* build_fit(const bvhar_fit_desc* D, int faithful, int rng_mode) {
  std::unique_ptr<Fit> F(new Fit);
  Shared& S = F->S;
  S.W = D->n_windows; S.C = D->n_chains; S.k = D->dim; S.d = D->dim_design;
  S.mean = D->include_mean != 0; S.sv = D->cov_type == BVHAR_COV_SV; S.ggl = D->is_group != 0;
  S.faithful = faithful != 0; S.philox = rng_mode == 0;
  S.prior = D->prior_type;
  S.num_iter = D->num_iter; S.num_burn = D->num_burn; S.thin = D->thin < 1 ? 1 : D->thin;
  S.rec_mask = D->record_mask; S.lvol_from_init = D->lvol_from_init != 0;
  S.q = S.k * (S.k - 1) / 2;
  const int num_coef = S.k * S.d;  // cfg.h:76-77
  S.N = S.mean ? num_coef - S.k : num_coef;
  S.nrow = S.N / S.k;
  S.G = D->n_grp;
  S.grp_vec.assign(D->grp_mat, D->grp_mat + S.N);
  S.grp_id.assign(D->grp_id, D->grp_id + S.G);
  for (int i = 0; i < D->n_own; ++i) S.own.insert(D->own_id[i]);
  S.sig_shp.assign(D->sig_shape, D->sig_shape + S.k);
  S.sig_scl.assign(D->sig_scale, D->sig_scale + S.k);
  if (S.sv) {
    S.sv_mean.assign(D->sv_init_mean, D->sv_init_mean + S.k);
    S.sv_prec = Mat(S.k, S.k);
    std::memcpy(S.sv_prec.data(), D->sv_init_prec, sizeof(double) * S.k * S.k);
  }
  if (S.mean) { S.mean_non.assign(D->mean_non, D->mean_non + S.k); S.sd_non = D->sd_non; }
  S.pp = D->prior;
  if (S.prior == BVHAR_PRIOR_MINNESOTA || S.prior == BVHAR_PRIOR_HIERMINN) {
    S.minn_mean.assign(D->prior.minn_prior_mean, D->prior.minn_prior_mean + S.N);
    S.minn_prec.assign(D->prior.minn_prior_prec, D->prior.minn_prior_prec + S.N);
  }
  if (S.prior == BVHAR_PRIOR_SSVS) {
    S.ssvs_s1.assign(D->prior.ssvs_coef_s1, D->prior.ssvs_coef_s1 + S.G);
    S.ssvs_s2.assign(D->prior.ssvs_coef_s2, D->prior.ssvs_coef_s2 + S.G);
  }
  F->wins.resize(S.W);
  const int64_t ld = D->n_rows_total;
  for (int w = 0; w < S.W; ++w) {
    Window& Wn = F->wins[w];
    Wn.row0 = D->win_row0[w]; Wn.n = D->win_nrow[w];
    Wn.X = Mat(Wn.n, S.d); Wn.Y = Mat(Wn.n, S.k);
    for (int a = 0; a < S.d; ++a) for (int t = 0; t < Wn.n; ++t) Wn.X(t, a) = D->x[(size_t)a * ld + Wn.row0 + t];
    for (int j = 0; j < S.k; ++j) for (int t = 0; t < Wn.n; ++t) Wn.Y(t, j) = D->y[(size_t)j * ld + Wn.row0 + t];
    Wn.xnorm.assign(S.d, 0.0);
    for (int a = 0; a < S.d; ++a) for (int t = 0; t < Wn.n; ++t) Wn.xnorm[a] += Wn.X(t, a) * Wn.X(t, a);
    gemm_tn(Wn.X, Wn.X, Wn.XtX);
    gemm_tn(Wn.X, Wn.Y, Wn.XtY);
  }
  F->chains.resize((size_t)S.W * S.C);
  for (int w = 0; w < S.W; ++w)
    for (int c = 0; c < S.C; ++c) {
      const int u = w * S.C + c;
      F->chains[u].reset(new Chain);
      F->chains[u]->init(&F->S, &F->wins[w], D->init[c], D->seed_chain[u], (uint32_t)(D->unit_id_base + u));
    }
  return F.release();
}

// McmcForecaster::forecastDensity + forecastOut (fc.h:73-85, 179-190), RegForecaster fc.h:206-222,
// SvForecaster fc.h:239-262, computeMean fc.h:300-302 / 338-340.
static void forecast_unit(const bvhar_forecast_desc* D, int u, bool philox, double* out, double* lpl_out) {
  const int k = D->dim, nrow = D->nrow_coef, S = D->n_draws, step = D->step, p = D->var_lag;
  const bool mean = D->include_mean != 0, sv_model = D->cov_type == BVHAR_COV_SV;
  const int ncoef = nrow * k + (mean ? k : 0), q = k * (k - 1) / 2;
  const int dim_design = p * k + (mean ? 1 : 0);
  const double* coef_rec = D->coef_record + (size_t)u * S * ncoef;
  const double* a_rec = D->contem_record + (size_t)u * S * q;
  const double* dg_rec = D->diag_record + (size_t)u * S * k;
  const double* sg_rec = sv_model ? D->sigh_record + (size_t)u * S * k : nullptr;
  const double* last = D->last_obs + (size_t)u * p * k;  // p x k col-major, oldest row first
  Rng rng;
  rng.seed(D->seed[u], (uint32_t)(D->unit_id_base + u), philox);
  Vec obs(dim_design, 0.0);  // [y_T, y_{T-1}, ..., 1]  fc.h:45-46
  if (mean) obs[dim_design - 1] = 1.0;
  for (int l = 0; l < p; ++l)
    for (int j = 0; j < k; ++j) obs[l * k + j] = last[(size_t)j * p + (p - 1 - l)];
  const int har_r = 3 * k + (mean ? 1 : 0), har_c = p * k + (mean ? 1 : 0);
  Vec lpl(step, 0.0), last_pvec, post_mean(k), tmp((p - 1) * k), svu(k), svs(k), zn(k), hx(D->is_vhar ? har_r : 0);
  Mat coef(nrow + (mean ? 1 : 0), k);
  // Select forecasters (fc.h:348-416): activity_graph = unvectorize(reg_record->computeActivity(level), dim), cfg.h:747-758.
  // quantile_lower / quantile_upper (common.h:585-606) are boost::accumulators::tail_quantile<left / right> with the whole
  // sample cached (Boost.Accumulators, BH >= 1.87, not vendored): n = ceil(count * prob) resp. ceil(count * (1. - prob)),
  // result = the n-th smallest resp. n-th largest sample, NaN unless n < count.
  Mat graph;
  if (D->level > 0) {
    const int df = nrow + (mean ? 1 : 0);
    graph = Mat(df, k);
    const double p_lo = D->level / 2, p_hi = 1 - D->level / 2;
    const size_t n_lo = (size_t)std::ceil((double)S * p_lo), n_hi = (size_t)std::ceil((double)S * (1. - p_hi));
    Vec col(S), sel(ncoef);
    for (int e = 0; e < ncoef; ++e) {
      for (int i = 0; i < S; ++i) col[i] = coef_rec[(size_t)e * S + i];
      std::sort(col.begin(), col.end());
      const double nan = std::numeric_limits<double>::quiet_NaN();
      const double lo = (n_lo >= 1 && n_lo < (size_t)S) ? col[n_lo - 1] : nan;
      const double hi = (n_hi >= 1 && n_hi < (size_t)S) ? col[S - n_hi] : nan;
      sel[e] = lo * hi < 0 ? 0.0 : 1.1;
    }
    for (int j = 0; j < k; ++j)
      for (int r = 0; r < df; ++r) graph(r, j) = sel[(size_t)j * df + r];  // column-major reshape of the VECTOR (fc.h:356)
  }
  for (int i = 0; i < S; ++i) {
    rng.iter = (uint32_t)i;
    last_pvec = obs;
    for (int j = 0; j < k; ++j) post_mean[j] = obs[j];
    for (int e = 0; e < (p - 1) * k; ++e) tmp[e] = obs[k + e];
    // updateParams
    for (int j = 0; j < k; ++j) {
      for (int r = 0; r < nrow; ++r) coef(r, j) = coef_rec[(size_t)(j * nrow + r) * S + i];
      if (mean) coef(nrow, j) = coef_rec[(size_t)(nrow * k + j) * S + i];
    }
    if (D->level > 0)  // computeMean of the Select variants: (activity_graph o coef_mat), fc.h:372-374, 409-411
      for (int j = 0; j < k; ++j)
        for (int r = 0; r < nrow + (mean ? 1 : 0); ++r) coef(r, j) *= graph(r, j);
    Vec a(q);
    for (int e = 0; e < q; ++e) a[e] = a_rec[(size_t)e * S + i];
    Mat L = build_inv_lower(k, a.data());
    for (int j = 0; j < k; ++j) {
      if (sv_model) { svu[j] = dg_rec[(size_t)j * S + i]; svs[j] = std::sqrt(sg_rec[(size_t)j * S + i]); }  // cfg.h:998-1001
      else svu[j] = std::sqrt(dg_rec[(size_t)j * S + i]);                                                 // cfg.h:876-878
    }
    for (int h = 0; h < step; ++h) {
      for (int e = 0; e < (p - 1) * k; ++e) last_pvec[k + e] = tmp[e];
      for (int j = 0; j < k; ++j) last_pvec[j] = post_mean[j];
      // computeMean
      if (D->is_vhar) {
        for (int r = 0; r < har_r; ++r) {
          double s = 0;
          for (int c = 0; c < har_c; ++c) s += D->har_trans[(size_t)c * har_r + r] * last_pvec[c];
          hx[r] = s;
        }
        for (int j = 0; j < k; ++j) { double s = 0; for (int r = 0; r < har_r; ++r) s += coef(r, j) * hx[r]; post_mean[j] = s; }
      } else {
        for (int j = 0; j < k; ++j) { double s = 0; for (int r = 0; r < dim_design; ++r) s += coef(r, j) * last_pvec[r]; post_mean[j] = s; }
      }
      // updateVariance
      if (sv_model) {
        if (D->sv) {
          rng.normal_vec(ST_FORECAST_H, (uint32_t)(h * k), k, zn.data());
          for (int j = 0; j < k; ++j) svu[j] += zn[j] * svs[j];
        }
        rng.normal_vec(ST_FORECAST_Y, (uint32_t)(h * k), k, zn.data());
        for (int j = 0; j < k; ++j) zn[j] *= std::exp(svu[j] / 2);
      } else {
        rng.normal_vec(ST_FORECAST_Y, (uint32_t)(h * k), k, zn.data());
        for (int j = 0; j < k; ++j) zn[j] *= svu[j];
      }
      // post_mean += L^-1 e (unit lower solve)
      for (int j = 0; j < k; ++j) { double s = zn[j]; for (int m = 0; m < j; ++m) s -= L(j, m) * zn[m]; zn[j] = s; }
      for (int j = 0; j < k; ++j) post_mean[j] += zn[j];
      for (int j = 0; j < k; ++j) out[(size_t)(i * k + j) * step + h] = post_mean[j];
      if (D->get_lpl && D->valid_vec) {  // fc.h:221 / 261 (sign conventions as written)
        double quad = 0, lg = 0;
        for (int j = 0; j < k; ++j) {
          double s = 0;
          for (int m = 0; m <= j; ++m) s += L(j, m) * (post_mean[m] - D->valid_vec[m]);
          if (sv_model) { s *= std::exp(-svu[j] / 2); lg += svu[j] / 2; }
          else { s /= svu[j]; lg += std::log(svu[j]); }
          quad += s * s;
        }
        lpl[h] += lg - k * std::log(2 * M_PI) / 2 - quad / 2;
      }
      for (int e = 0; e < (p - 1) * k; ++e) tmp[e] = last_pvec[e];
    }
  }
  if (lpl_out) for (int h = 0; h < step; ++h) lpl_out[h] = lpl[h] / S;
}

}  // namespace orc

using namespace orc;

extern "C" {

void* oracle_fit_create(const bvhar_fit_desc* desc, int faithful, int rng_mode) { return build_fit(desc, faithful, rng_mode); }
void oracle_fit_destroy(void* h) { delete (Fit*)h; }
void oracle_fit_set_threads(void* h, int nt) { ((Fit*)h)->nthreads = nt; }

int oracle_fit_step(void* h, int n_sweeps, int record) {
  Fit* F = (Fit*)h;
  const int U = (int)F->chains.size();
  const int nt = F->nthreads > 0 ? F->nthreads : omp_get_max_threads();
  const int outer = U < nt ? U : nt, inner = U < nt ? nt / U : 1;
  omp_set_max_active_levels(2);
  for (auto& c : F->chains) c->inner_threads = inner;
#pragma omp parallel for num_threads(outer) schedule(dynamic, 1)   // tri.h:1222 / fc.h:626: chains and windows
  for (int u = 0; u < U; ++u)
    for (int s = 0; s < n_sweeps; ++s) F->chains[u]->sweep(record != 0);
  return 0;
}
// runGibbs tri.h:1241-1284
int oracle_fit_run(void* h) {
  Fit* F = (Fit*)h;
  oracle_fit_step(h, F->S.num_burn, 0);
  oracle_fit_step(h, F->S.num_iter - F->S.num_burn, 1);
  return 0;
}
int oracle_fit_run_stage(void* h, int stage, int iter) {
  Fit* F = (Fit*)h;
  for (auto& c : F->chains) { c->rng.iter = (uint32_t)iter; c->stage(stage); }
  return 0;
}
int64_t oracle_fit_state_len(void* h, const char* name) {
  Fit* F = (Fit*)h;
  return (int64_t)F->chains[0]->state(name).len;
}
int oracle_fit_get_state(void* h, int w, int c, const char* name, double* out, int64_t len) {
  Fit* F = (Fit*)h;
  Chain::Ref r = F->chains[(size_t)w * F->S.C + c]->state(name);
  if (!r.p || (int64_t)r.len != len) return 1;
  std::memcpy(out, r.p, sizeof(double) * len);
  return 0;
}
int oracle_fit_set_state(void* h, int w, int c, const char* name, const double* in, int64_t len) {
  Fit* F = (Fit*)h;
  Chain::Ref r = F->chains[(size_t)w * F->S.C + c]->state(name);
  if (!r.p || (int64_t)r.len != len) return 1;
  std::memcpy(r.p, in, sizeof(double) * len);
  return 0;
}
int oracle_fit_set_sweep_count(void* h, int count) {
  for (auto& c : ((Fit*)h)->chains) c->sweep_count = (uint32_t)count;
  return 0;
}
// thin_record draw.h:1297-1310 with (num_iter = kept, num_burn = 0): rows 0, thin, 2 thin, ...
int oracle_fit_record_dims(void* h, const char* name, int64_t* rows, int64_t* cols) {
  Fit* F = (Fit*)h;
  Records& R = F->chains[0]->rec;
  auto it = R.m.find(name);
  if (it == R.m.end()) return 1;
  const int filled = R.filled, thin = F->S.thin;
  *rows = (filled + thin - 1) / thin;
  *cols = it->second.c;
  return 0;
}
int oracle_fit_record(void* h, int w, int c, const char* name, double* out, int64_t rows, int64_t cols) {
  Fit* F = (Fit*)h;
  Records& R = F->chains[(size_t)w * F->S.C + c]->rec;
  auto it = R.m.find(name);
  if (it == R.m.end()) return 1;
  const Mat& M = it->second;
  const int thin = F->S.thin;
  if (cols != M.c || rows != (R.filled + thin - 1) / thin) return 2;
  for (int64_t j = 0; j < cols; ++j)
    for (int64_t r = 0; r < rows; ++r) out[j * rows + r] = M((int)(r * thin), (int)j);
  return 0;
}
int oracle_fit_record_names(void* h, char* buf, int64_t buflen) {
  Fit* F = (Fit*)h;
  std::string s;
  for (auto& nm : F->chains[0]->rec.order) { s += nm; s += '\n'; }
  if ((int64_t)s.size() + 1 > buflen) return 1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

int oracle_forecast_density(const bvhar_forecast_desc* D, int rng_mode, double* out, double* lpl) {
  const size_t per = (size_t)D->step * D->n_draws * D->dim;
#pragma omp parallel for schedule(dynamic, 1)   // fc.h:545
  for (int u = 0; u < D->n_units; ++u)
    forecast_unit(D, u, rng_mode == 0, out + per * u, lpl ? lpl + (size_t)D->step * u : nullptr);
  return 0;
}

// ---- deterministic stage functions with the random numbers passed in ----------------------
// draw_coef draw.h:372-383: x (n x d), y (n), prior mean / precision (d), z (d) -> coef (d)
void oracle_draw_coef(int n, int d, const double* x, const double* y, const double* prior_mean,
                      const double* prior_prec, const double* z, double* coef) {
  Mat X(n, d);
  std::memcpy(X.data(), x, sizeof(double) * n * d);
  Mat P; gemm_tn(X, X, P);
  Vec rhs(d);
  for (int a = 0; a < d; ++a) {
    P(a, a) += prior_prec[a];
    double s = 0;
    for (int t = 0; t < n; ++t) s += X(t, a) * y[t];
    rhs[a] = prior_prec[a] * prior_mean[a] + s;
  }
  precision_draw(P, rhs.data(), z, coef);
}
// packed-lower precision in, as the CUDA building block takes it
void oracle_draw_coef_packed(int d, const double* P_packed, const double* b, const double* z, double* coef) {
  Mat P(d, d);
  for (int a = 0; a < d; ++a)
    for (int c = 0; c <= a; ++c) P(a, c) = P_packed[(size_t)a * (a + 1) / 2 + c];
  precision_draw(P, b, z, coef);
}
void oracle_dgemm_tn(int M, int N, int K, const double* A, const double* B, double* C) {
  // A: K x M row-major, B: K x N row-major, C: M x N row-major
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int t = 0; t < K; ++t) s += A[(size_t)t * M + i] * B[(size_t)t * N + j];
      C[(size_t)i * N + j] = s;
    }
}
// ---- spillover of one draw: McmcSpillover::computeSpillover (spillover.h:48-66), written as the reference writes it ----
// convert_var_to_vma / convert_vhar_to_vma (math/structural.h:8-36, 54-78)
static Mat oracle_var_to_vma(const Mat& var_coef, int var_lag, int lag_max) {
  const int dim = var_coef.c;
  const int ma_rows = dim * (lag_max + 1);
  int num_full_arows = ma_rows;
  if (lag_max < var_lag) num_full_arows = dim * var_lag;
  Mat FullA(num_full_arows, dim);
  for (int j = 0; j < dim; ++j)
    for (int i = 0; i < dim * var_lag; ++i) FullA(i, j) = var_coef(i, j);
  Mat ma(ma_rows, dim);
  for (int i = 0; i < dim; ++i) ma(i, i) = 1.0;  // W0 = Im
  if (lag_max >= 1)
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < dim; ++c) {
        double sacc = 0;
        for (int m = 0; m < dim; ++m) sacc += FullA(r, m) * ma(m, c);
        ma(dim + r, c) = sacc;
      }
  for (int i = 2; i < lag_max + 1; ++i)
    for (int kq = 0; kq < i; ++kq)
      for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
          double sacc = 0;
          for (int m = 0; m < dim; ++m) sacc += FullA(kq * dim + r, m) * ma((i - kq - 1) * dim + m, c);
          ma(i * dim + r, c) += sacc;
        }
  return ma;
}
// compute_vma_fevd (math/structural.h:100-127)
static Mat oracle_vma_fevd(const Mat& vma, const Mat& cov, bool normalize) {
  const int dim = cov.c, step = vma.r / dim;
  Mat innov(dim, dim), numer(dim, dim), res(dim * step, dim), ma_prod(dim, dim);
  for (int i = 0; i < step; ++i) {
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < dim; ++c) {
        double sacc = 0;
        for (int m = 0; m < dim; ++m) sacc += vma(i * dim + m, r) * cov(m, c);  // block' * Sigma
        ma_prod(r, c) = sacc;
      }
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < dim; ++c) {
        double sacc = 0;
        for (int m = 0; m < dim; ++m) sacc += ma_prod(r, m) * vma(i * dim + m, c);
        innov(r, c) += sacc;
        const double v = ma_prod(r, c) * (1.0 / std::sqrt(cov(c, c)));
        numer(r, c) += v * v;
      }
    for (int r = 0; r < dim; ++r)
      for (int c = 0; c < dim; ++c) res(i * dim + r, c) = (1.0 / innov(r, r)) * numer(r, c);
  }
  if (normalize)
    for (int r = 0; r < dim * step; ++r) {
      double rs = 0;
      for (int c = 0; c < dim; ++c) rs += res(r, c);
      for (int c = 0; c < dim; ++c) res(r, c) /= rs;
    }
  return res;
}
// alpha: equation-major coefficients of one draw; a: q contemporaneous coefficients; sv: D^(1/2) (dim); har: 3 dim x
// (lag dim) column-major or NULL.  Outputs: connect / net_pairwise dim x dim column-major, to / from dim, tot 1.
void oracle_spillover(int dim, int nrow, int lag, int step, const double* alpha, const double* a, const double* sv, const double* har,
                      double* connect, double* to, double* from, double* tot, double* net_pairwise) {
  Mat L(dim, dim);  // build_inv_lower (draw.h:124-132)
  for (int i = 0; i < dim; ++i) {
    L(i, i) = 1.0;
    for (int j = 0; j < i; ++j) L(i, j) = a[i * (i - 1) / 2 + j];
  }
  Mat sq(dim, dim);  // L^-1 diag(sv): unit-lower solve, column by column
  for (int c = 0; c < dim; ++c)
    for (int r = 0; r < dim; ++r) {
      double sacc = (r == c) ? sv[c] : 0.0;
      for (int m = 0; m < r; ++m) sacc -= L(r, m) * sq(m, c);
      sq(r, c) = sacc;
    }
  Mat cov(dim, dim);
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) {
      double sacc = 0;
      for (int m = 0; m < dim; ++m) sacc += sq(r, m) * sq(c, m);
      cov(r, c) = sacc;
    }
  Mat coef(nrow, dim);  // unvectorize
  for (int j = 0; j < dim; ++j)
    for (int i = 0; i < nrow; ++i) coef(i, j) = alpha[(size_t)j * nrow + i];
  Mat var_coef = coef;
  if (har) {  // HARtrans' * Phi (structural.h:56)
    var_coef = Mat(lag * dim, dim);
    for (int r = 0; r < lag * dim; ++r)
      for (int j = 0; j < dim; ++j) {
        double sacc = 0;
        for (int qq = 0; qq < nrow; ++qq) sacc += har[(size_t)r * nrow + qq] * coef(qq, j);
        var_coef(r, j) = sacc;
      }
  }
  const Mat vma = oracle_var_to_vma(var_coef, lag, step - 1);
  const Mat fevd = oracle_vma_fevd(vma, cov, true);
  Mat sp(dim, dim);  // compute_sp_index: bottom rows * 100
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) sp(r, c) = fevd((step - 1) * dim + r, c) * 100;
  double off = 0;
  for (int c = 0; c < dim; ++c) {
    double tsum = 0, fsum = 0;
    for (int r = 0; r < dim; ++r) {
      connect[(size_t)c * dim + r] = sp(r, c);
      net_pairwise[(size_t)c * dim + r] = (sp(c, r) - sp(r, c)) / dim;
      if (r != c) { tsum += sp(r, c); fsum += sp(c, r); off += sp(r, c); }
    }
    to[c] = tsum;
    from[c] = fsum;
  }
  *tot = off / dim;
}
void oracle_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// Sampler access for distribution tests. kind: 0 normal, 1 uniform, 2 gamma(p0,p1), 3 beta(p0,p1),
// 4 invgauss(p0,p1), 5 gig(p0,p1,p2).  One draw per element index, iteration `iter`.
void oracle_sample(int kind, int rng_mode, uint32_t seed, uint32_t iter, double p0, double p1, double p2, int n, double* out) {
  Rng rng;
  rng.seed(seed, 0, rng_mode == 0);
  rng.iter = iter;
  for (int i = 0; i < n; ++i) {
    switch (kind) {
      case 0: rng.normal_vec(ST_PRIOR_COEF, (uint32_t)i, 1, out + i); break;
      case 1: rng.uniform_vec(ST_PRIOR_COEF, (uint32_t)i, 1, out + i); break;
      case 2: out[i] = rng.gamma(ST_PRIOR_COEF, (uint32_t)i, p0, p1); break;
      case 3: out[i] = rng.beta(ST_PRIOR_COEF, (uint32_t)i, p0, p1); break;
      case 4: out[i] = rng.invgauss(ST_PRIOR_COEF, (uint32_t)i, p0, p1); break;
      case 5: out[i] = rng.gig(ST_PRIOR_COEF, (uint32_t)i, p0, p1, p2); break;
    }
  }
}

// One variate of the addressed Philox stream (DESIGN.md "Random streams"): sampler `kind` at (seed, unit | iter, stage,
// idx).  kind: 0 normal element idx of a vector site, 1 uniform element idx of a vector site, 2 gamma(p0 shape, p1 scale),
// 3 beta(p0, p1), 4 invgauss(p0 mean, p1 shape), 5 gig(p0, p1, p2), 6 scalar-site uniform (Bernoulli / categorical
// sites compare it themselves).  Lets an independent restatement of the sweep (tests/numpy_mirror.py) be fed the very
// same random variates as the oracle, so that only the algebra is under test.
double oracle_draw_at(int kind, uint32_t seed, uint32_t unit, uint32_t iter, uint32_t stage, uint32_t idx, double p0, double p1, double p2) {
  Rng rng;
  rng.seed(seed, unit, true);
  rng.iter = iter;
  double out = 0.0;
  switch (kind) {
    case 0: rng.normal_vec(stage, idx, 1, &out); return out;
    case 1: rng.uniform_vec(stage, idx, 1, &out); return out;
    case 2: return rng.gamma(stage, idx, p0, p1);
    case 3: return rng.beta(stage, idx, p0, p1);
    case 4: return rng.invgauss(stage, idx, p0, p1);
    case 5: return rng.gig(stage, idx, p0, p1, p2);
    case 6: return rng.uniform(stage, idx, 0);
  }
  return std::numeric_limits<double>::quiet_NaN();
}

// Minnesota prior moments: build_ydummy / build_xdummy design.h:60-98 and cfg.h:144-151,
// prior_alpha_prec = diag(kron(diag(1/sigma), Xd'Xd)) tri.h:438-439.  `lag` = p (3 for VHAR).
void oracle_minnesota_moments(int k, int lag, const double* sigma, double lambda, double eps, const double* daily,
                              const double* weekly, const double* monthly, double* prior_mean /* (k*lag) x k */,
                              double* prior_prec_full /* (k*lag) x (k*lag) */, double* alpha_prec /* k*k*lag */) {
  const int nr = k * lag + k, nc = k * lag;  // include_mean = false variants
  Mat Yd(nr, k), Xd(nr, nc);
  for (int i = 0; i < k; ++i) {
    Yd(i, i) = daily[i] * sigma[i] / lambda;
    if (lag > 1) { Yd(k + i, i) = weekly[i] * sigma[i] / lambda; Yd(2 * k + i, i) = monthly[i] * sigma[i] / lambda; }
    Yd(k * lag + i, i) = sigma[i];
  }
  for (int l = 0; l < lag; ++l)
    for (int i = 0; i < k; ++i) Xd(l * k + i, l * k + i) = (l + 1) * sigma[i] / lambda;
  (void)eps;  // only enters the constant-term row/column, dropped when include_mean = false (design.h:93-97)
  Mat P, XY;
  gemm_tn(Xd, Xd, P);
  gemm_tn(Xd, Yd, XY);
  std::memcpy(prior_prec_full, P.data(), sizeof(double) * nc * nc);
  Mat Lc = P;
  chol_lower(Lc);
  for (int j = 0; j < k; ++j) {
    Vec col(XY.col(j), XY.col(j) + nc);
    solve_lower(Lc, col.data());
    solve_lower_t(Lc, col.data());
    for (int i = 0; i < nc; ++i) prior_mean[(size_t)j * nc + i] = col[i];
  }
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < nc; ++i) alpha_prec[(size_t)j * nc + i] = (1.0 / sigma[j]) * P(i, i);
}

int oracle_max_threads() { return omp_get_max_threads(); }

// Route the large dense products / Cholesky factorisations through a BLAS shared library (bench.py's reference arm: the
// OpenBLAS bundled with SciPy, symbols scipy_cblas_dgemm / scipy_dpotrf_ / scipy_openblas_set_num_threads).  The library is
// put in single-thread mode: the parallelism stays the reference's (one OpenMP task per chain, tri.h:1222).  Returns 0 on
// success; path == NULL switches back to the hand-written loops.
int oracle_use_blas(const char* path, const char* prefix) {
  if (!path) { blas_dgemm() = nullptr; lapack_dpotrf() = nullptr; return 0; }
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  const std::string pre = prefix ? prefix : "";
  auto sym = [&](const char* nm) { return dlsym(h, (pre + nm).c_str()); };
  void* g = sym("cblas_dgemm");
  void* c = sym("dpotrf_");
  void* t = sym("openblas_set_num_threads");
  if (!g || !c) return 2;
  if (t) ((void (*)(int))t)(1);
  blas_dgemm() = (cblas_dgemm_fn)g;
  lapack_dpotrf() = (dpotrf_fn)c;
  return 0;
}

// ---- conjugate Minnesota samplers: McmcMniw / MhMinnesota (bayes/mniw/minnesota.h:241-278, 413-510) ----------------
// A literal restatement of one chain's iteration loop with the draws addressed on the Philox stream
// (stages ST_MNIW_* / ST_MH_*, iteration = mcmc_step):
//   sim_iw_tri  (math/random.h:177-200)  upper Bartlett factor Q: Q_ii = sqrt(chi^2(shape - i)), Q_ij ~ N(0,1) (j > i);
//                                        result = llt(scale).L * (Q' lower-triangular)^-1
//   sim_mn_iw   (random.h:203-211)       Sigma = res res'; coef = sim_mn(mean, prec, Sigma, prec = true)
//   sim_mn      (random.h:158-175)       mean + llt(prec).U^-1 (Z * llt(Sigma).U),  Z filled row by row
//   updateHyper (minnesota.h:446-461)    candprior = prevprior + z' llt(gaussian_variance).U  (sim_mgaussian_chol,
//                                        random.h:9-21); prevprior is never moved (reference behaviour);
//                                        accept = log(u) < min(numerator - denom, 0).  compute_logml and the constant
//                                        terms of jointdens_hyperparam (draw.h:59-72) take identical arguments in the
//                                        numerator and the denominator and cancel; gamma_shp / gamma_rate are swapped as
//                                        in minnesota.h:426.
// Records: row r of the thinned record is mcmc_step num_burn + 1 + r * thin (thin_record, draw.h:1297-1310); outputs are
// column-major rows x cols per chain.  Returns 3 when a proposed psi is negative (invgamma_dens STOPs, common.h:503).
static double orc_gamma_logdens(double x, double shp, double scl) {
  if (!(x > 0.0)) return -std::numeric_limits<double>::infinity();
  return (shp - 1.0) * std::log(x) - x / scl - std::lgamma(shp) - shp * std::log(scl);
}
static double orc_invgamma_logdens(double x, double shp, double scl) {
  return std::log(std::pow(scl, shp) * std::pow(x, -shp - 1.0) * std::exp(-scl / x) / std::tgamma(shp));
}

int oracle_mniw_run(int num_chains, int num_iter, int num_burn, int thin, int k, int d, const double* mn_mean, const double* mn_prec,
                    const double* iw_scale, double iw_shape, const uint32_t* seed_chain, uint32_t unit_base, int mh, double gam_shape,
                    double gam_rate, double invgam_shape, double invgam_scl, const double* init_par, const double* chol_var_upper,
                    double* alpha_record, double* sigma_record, double* lambda_record, double* psi_record, uint8_t* accept_record) {
  if (thin < 1) thin = 1;
  const int rows = (num_iter - num_burn + thin - 1) / thin;
  Mat Lu(d, d), Ls(k, k);
  for (int j = 0; j < d; ++j) for (int i = 0; i < d; ++i) Lu(i, j) = mn_prec[(size_t)j * d + i];
  for (int j = 0; j < k; ++j) for (int i = 0; i < k; ++i) Ls(i, j) = iw_scale[(size_t)j * k + i];
  if (!chol_lower(Lu) || !chol_lower(Ls)) return 2;
  const double gamma_shp = gam_rate, gamma_rate = gam_shape;  // minnesota.h:426
  const int np = 1 + k;
  int status = 0;
  for (int c = 0; c < num_chains; ++c) {
    Rng rng;
    rng.seed(seed_chain[c], unit_base + (uint32_t)c, true);
    Vec lam_psi(np, 0.0);
    if (mh) for (int j = 0; j < np; ++j) lam_psi[j] = init_par[(size_t)c * np + j];
    bool is_accept = true;
    for (int t = 1; t <= num_iter; ++t) {
      rng.iter = (uint32_t)t;
      if (mh) {
        const double* prev = init_par + (size_t)c * np;
        const double* G = chol_var_upper + (size_t)c * np * np;  // row-major upper factor
        Vec z(np), cand(np);
        rng.normal_vec(ST_MH_CAND, 0, np, z.data());
        for (int j = 0; j < np; ++j) {
          double v = prev[j];
          for (int i = 0; i <= j; ++i) v = std::fma(z[i], G[(size_t)i * np + j], v);
          cand[j] = v;
        }
        double num = orc_gamma_logdens(cand[0], gamma_shp, 1.0 / gamma_rate), den = orc_gamma_logdens(prev[0], gamma_shp, 1.0 / gamma_rate);
        for (int j = 1; j < np; ++j) {
          if (cand[j] < 0.0) status = 3;
          num += orc_invgamma_logdens(cand[j], invgam_shape, invgam_scl);
          den += orc_invgamma_logdens(prev[j], invgam_shape, invgam_scl);
        }
        const double u = rng.uniform(ST_MH_U, 0, 0);
        is_accept = std::log(u) < std::min(num - den, 0.0);
        if (is_accept) lam_psi = cand;
      }
      const int post = t - num_burn - 1;
      if (post < 0 || post % thin) continue;  // only the kept iterations are materialised (draws are addressed)
      const int r = post / thin;
      // sim_iw_tri
      Mat Q(k, k);
      for (int i = 0; i < k; ++i) Q(i, i) = std::sqrt(rng.gamma(ST_MNIW_CHI, (uint32_t)i, 0.5 * (iw_shape - (double)i), 2.0));
      for (int i = 0; i < k - 1; ++i)
        for (int j = i + 1; j < k; ++j) {
          const uint32_t e = (uint32_t)(i * k + j);
          double z0, z1;
          rng.normal_pair(ST_MNIW_BART, e >> 1, 0, z0, z1);
          Q(i, j) = (e & 1u) ? z1 : z0;
        }
      Mat Binv(k, k);  // (Q')^-1, column by column: solve Q' y = e_c
      for (int cc = 0; cc < k; ++cc) {
        Vec y(k, 0.0);
        y[cc] = 1.0;
        for (int i = 0; i < k; ++i) {
          double sacc = y[i];
          for (int m = 0; m < i; ++m) sacc -= Q(m, i) * y[m];
          y[i] = sacc / Q(i, i);
        }
        for (int i = 0; i < k; ++i) Binv(i, cc) = y[i];
      }
      Mat Cm(k, k), Sig(k, k);
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
          double sacc = 0.0;
          for (int m = 0; m < k; ++m) sacc += (m <= i ? Ls(i, m) : 0.0) * Binv(m, j);
          Cm(i, j) = sacc;
        }
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
          double sacc = 0.0;
          for (int m = 0; m < k; ++m) sacc += Cm(i, m) * Cm(j, m);
          Sig(i, j) = sacc;
        }
      for (int j = 0; j < k; ++j)
        for (int i = 0; i < k; ++i) sigma_record[((size_t)c * k * k + (size_t)j * k + i) * rows + r] = Sig(i, j);
      // sim_mn with prec = true
      Mat Lv = Sig;
      if (!chol_lower(Lv)) status = status ? status : 2;
      Mat Z(d, k);
      for (int i = 0; i < d; ++i) {
        Vec row(k);
        rng.normal_vec(ST_MNIW_Z, (uint32_t)(i * k), k, row.data());
        for (int j = 0; j < k; ++j) Z(i, j) = row[j];
      }
      Mat W(d, k);
      for (int i = 0; i < d; ++i)
        for (int cc = 0; cc < k; ++cc) {
          double sacc = 0.0;
          for (int j = 0; j <= cc; ++j) sacc += Z(i, j) * Lv(cc, j);  // U_v(j, cc) = L_v(cc, j)
          W(i, cc) = sacc;
        }
      for (int cc = 0; cc < k; ++cc) {
        solve_lower_t(Lu, W.col(cc));  // U_u x = w with U_u = L_u'
        for (int i = 0; i < d; ++i)
          alpha_record[((size_t)c * d * k + (size_t)cc * d + i) * rows + r] = mn_mean[(size_t)cc * d + i] + W(i, cc);
      }
      if (mh) {
        lambda_record[(size_t)c * rows + r] = lam_psi[0];
        for (int j = 1; j < np; ++j) psi_record[((size_t)c * k + (j - 1)) * rows + r] = lam_psi[j];
        accept_record[(size_t)c * rows + r] = is_accept ? 1 : 0;
      }
    }
  }
  return status;
}

}  // extern "C"
