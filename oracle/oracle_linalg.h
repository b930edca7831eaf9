// TEST INFRASTRUCTURE ONLY (see oracle_rng.h).  Small dense column-major linear algebra for the
// oracle: stands in for the Eigen 3.4 calls of the reference (LLT, triangularView::solve, GEMM,
// kroneckerProduct; un-vendored dependency RcppEigen >= 0.3.4, DESCRIPTION:46).  Hand-written
// loops; an optional BLAS back end for the big faithful-mode products (see blas_dgemm below).
#ifndef BVHAR_ORACLE_LINALG_H
#define BVHAR_ORACLE_LINALG_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace orc {

struct Mat {
  int r = 0, c = 0;
  std::vector<double> v;
  Mat() {}
  Mat(int r_, int c_, double fill = 0.0) : r(r_), c(c_), v((size_t)r_ * c_, fill) {}
  double& operator()(int i, int j) { return v[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return v[(size_t)j * r + i]; }
  double* col(int j) { return v.data() + (size_t)j * r; }
  const double* col(int j) const { return v.data() + (size_t)j * r; }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  size_t size() const { return v.size(); }
};
typedef std::vector<double> Vec;

// Optional BLAS / LAPACK back end for the LARGE dense products and factorisations of the reference-faithful formulation
// (the 1000 x 1000 H'H product and LLT of varsv_ht, the Kronecker-design Gram): the reference runs them through Eigen's
// vectorised kernels, so timing the oracle with scalar loops would flatter the GPU.  bench.py's reference arm points these
// at the OpenBLAS that ships inside SciPy (oracle_use_blas, bvhar_oracle.cpp); the tests never do, so that the golden
// trajectories do not depend on a third-party library's rounding.
typedef void (*cblas_dgemm_fn)(int order, int transa, int transb, int m, int n, int k, double alpha, const double* a, int lda,
                               const double* b, int ldb, double beta, double* c, int ldc);
typedef void (*dpotrf_fn)(const char* uplo, const int* n, double* a, const int* lda, int* info);
inline cblas_dgemm_fn& blas_dgemm() { static cblas_dgemm_fn f = nullptr; return f; }
inline dpotrf_fn& lapack_dpotrf() { static dpotrf_fn f = nullptr; return f; }

// C (m x n) = A^T B, A (k x m), B (k x n); cache-blocked over k, inner loop contiguous in k.
inline void gemm_tn(const Mat& A, const Mat& B, Mat& C) {
  const int m = A.c, n = B.c, k = A.r;
  C = Mat(m, n);
  if (blas_dgemm() && (double)m * n * k >= 262144.0) {  // CblasColMajor = 102, CblasTrans = 112, CblasNoTrans = 111
    blas_dgemm()(102, 112, 111, m, n, k, 1.0, A.data(), k, B.data(), k, 0.0, C.data(), m);
    return;
  }
  const int KB = 256;
  for (int k0 = 0; k0 < k; k0 += KB) {
    const int k1 = std::min(k, k0 + KB);
    for (int j = 0; j < n; ++j) {
      const double* b = B.col(j);
      for (int i = 0; i < m; ++i) {
        const double* a = A.col(i);
        double s = 0;
#pragma omp simd reduction(+ : s)
        for (int t = k0; t < k1; ++t) s += a[t] * b[t];
        C(i, j) += s;
      }
    }
  }
}
// C (m x n) = A B, A (m x k), B (k x n); axpy form, contiguous in m.
inline void gemm_nn(const Mat& A, const Mat& B, Mat& C) {
  const int m = A.r, n = B.c, k = A.c;
  C = Mat(m, n);
  if (blas_dgemm() && (double)m * n * k >= 262144.0) {
    blas_dgemm()(102, 111, 111, m, n, k, 1.0, A.data(), m, B.data(), k, 0.0, C.data(), m);
    return;
  }
  for (int j = 0; j < n; ++j) {
    double* c = C.col(j);
    for (int t = 0; t < k; ++t) {
      const double b = B(t, j);
      if (b == 0.0) continue;
      const double* a = A.col(t);
      for (int i = 0; i < m; ++i) c[i] += a[i] * b;
    }
  }
}

// In-place lower Cholesky of the lower triangle of P (d x d col-major) — Eigen::LLT semantics
// (only the lower triangle is referenced, draw.h:380).  Returns false if a pivot is not positive
// (Eigen would silently produce NaNs; the oracle keeps going with NaNs as well).
inline bool chol_lower(Mat& P) {
  const int d = P.r;
  bool ok = true;
  if (lapack_dpotrf() && d >= 96) {
    int info = 0;
    lapack_dpotrf()("L", &d, P.data(), &d, &info);
    return info == 0;
  }
  for (int j = 0; j < d; ++j) {
    double* cj = P.col(j);
    for (int t = 0; t < j; ++t) {
      const double ljt = P(j, t);
      if (ljt == 0.0) continue;
      const double* ct = P.col(t);
      for (int i = j; i < d; ++i) cj[i] -= ct[i] * ljt;
    }
    double piv = cj[j];
    if (!(piv > 0.0)) ok = false;
    piv = std::sqrt(piv);
    cj[j] = piv;
    const double inv = 1.0 / piv;
    for (int i = j + 1; i < d; ++i) cj[i] *= inv;
  }
  return ok;
}
// Solve L x = b in place.
inline void solve_lower(const Mat& L, double* x) {
  const int d = L.r;
  for (int j = 0; j < d; ++j) {
    x[j] /= L(j, j);
    const double xj = x[j];
    const double* c = L.col(j);
    for (int i = j + 1; i < d; ++i) x[i] -= c[i] * xj;
  }
}
// Solve L^T x = b in place.
inline void solve_lower_t(const Mat& L, double* x) {
  const int d = L.r;
  for (int j = d - 1; j >= 0; --j) {
    const double* c = L.col(j);
    double s = x[j];
    for (int i = j + 1; i < d; ++i) s -= c[i] * x[i];
    x[j] = s / c[j];
  }
}

// draw.h:372-383 with the standard normals passed in: P = diag(prec) + G (G lower valid),
// coef = P^-1 rhs + chol(P)^-T z.
inline void precision_draw(Mat& P, const double* rhs, const double* z, double* coef) {
  const int d = P.r;
  chol_lower(P);
  Vec m(rhs, rhs + d), e(z, z + d);
  solve_lower(P, m.data());    // llt.solve(): L y = rhs ...
  solve_lower_t(P, m.data());  // ... L^T mean = y                (draw.h:382)
  solve_lower_t(P, e.data());  // llt.matrixU().solve(res)        (draw.h:383)
  for (int i = 0; i < d; ++i) coef[i] = m[i] + e[i];
}

// draw.h:124-132
inline Mat build_inv_lower(int dim, const double* a) {
  Mat L(dim, dim);
  for (int i = 0; i < dim; ++i) L(i, i) = 1.0;
  int id = 0;
  for (int i = 1; i < dim; ++i) {
    for (int j = 0; j < i; ++j) L(i, j) = a[id + j];
    id += i;
  }
  return L;
}

}  // namespace orc
#endif
