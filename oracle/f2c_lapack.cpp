// TEST INFRASTRUCTURE ONLY -- the three LAPACK routines the reference's hot path calls
// (DGETRF src/conhoms.F:627, DGETRS src/ocisubs.F:359, DGERFS src/ocisubs.F:368).  LAPACK is a
// third-party dependency the reference does not vendor (src/lapack/ is absent, no version is
// pinned: SURVEY.md 8c), so the published algorithms are restated here with the LAPACK
// calling convention (column-major, by reference) and linked to the translated reference:
//   DGETRF: LU with partial pivoting (unblocked, as DGETF2; first maximum on ties like IDAMAX)
//   DGETRS: row interchanges, unit-lower then upper triangular solves ('N' only)
//   DGERFS: iterative refinement on the residual with the componentwise backward error
//           stopping rule of the LAPACK source (eps, ITMAX = 5); the error bounds FERR are not
//           estimated (the reference never reads them) and are returned as zero.
#include <cmath>
#include <limits>
#include <string>
#include <vector>

void dgetrf_(int *m_, int *n_, double *a, int *lda_, int *ipiv, int *info) {
  const int m = *m_, n = *n_, lda = *lda_;
  *info = 0;
  for (int j = 0; j < std::min(m, n); ++j) {
    int p = j;
    for (int i = j + 1; i < m; ++i)
      if (std::fabs(a[i + (long)lda * j]) > std::fabs(a[p + (long)lda * j])) p = i;
    ipiv[j] = p + 1;
    if (a[p + (long)lda * j] != 0.0) {
      if (p != j)
        for (int k = 0; k < n; ++k) std::swap(a[j + (long)lda * k], a[p + (long)lda * k]);
      const double r = 1.0 / a[j + (long)lda * j];
      for (int i = j + 1; i < m; ++i) a[i + (long)lda * j] *= r;
    } else if (*info == 0) {
      *info = j + 1;
    }
    for (int k = j + 1; k < n; ++k)
      for (int i = j + 1; i < m; ++i) a[i + (long)lda * k] -= a[i + (long)lda * j] * a[j + (long)lda * k];
  }
}

void dgetrs_(std::string trans, int *n_, int *nrhs_, double *a, int *lda_, int *ipiv, double *b, int *ldb_, int *info) {
  const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
  *info = (trans.empty() || (trans[0] != 'N' && trans[0] != 'n')) ? -1 : 0;
  if (*info) return;
  for (int c = 0; c < nrhs; ++c) {
    double *x = b + (long)ldb * c;
    for (int i = 0; i < n; ++i)
      if (ipiv[i] - 1 != i) std::swap(x[i], x[ipiv[i] - 1]);
    for (int j = 0; j < n; ++j)
      for (int i = j + 1; i < n; ++i) x[i] -= a[i + (long)lda * j] * x[j];
    for (int j = n - 1; j >= 0; --j) {
      x[j] /= a[j + (long)lda * j];
      for (int i = 0; i < j; ++i) x[i] -= a[i + (long)lda * j] * x[j];
    }
  }
}

void dgerfs_(std::string trans, int *n_, int *nrhs_, double *a, int *lda_, double *af, int *ldaf_, int *ipiv, double *b, int *ldb_,
             double *x, int *ldx_, double *ferr, double *berr, double *work, int *iwork, int *info) {
  (void)work; (void)iwork;
  const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_, ldx = *ldx_;
  *info = 0;
  const double eps = std::numeric_limits<double>::epsilon() * 0.5, safmin = std::numeric_limits<double>::min();
  const double safe1 = (n + 1) * safmin, safe2 = safe1 / eps;
  std::vector<double> r(n), w(n);
  for (int c = 0; c < nrhs; ++c) {
    double *xc = x + (long)ldx * c;
    const double *bc = b + (long)ldb * c;
    double lstres = 3.0;
    for (int count = 1;; ++count) {
      for (int i = 0; i < n; ++i) { r[i] = bc[i]; w[i] = std::fabs(bc[i]); }
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          r[i] -= a[i + (long)lda * j] * xc[j];
          w[i] += std::fabs(a[i + (long)lda * j]) * std::fabs(xc[j]);
        }
      double s = 0.0;
      for (int i = 0; i < n; ++i)
        s = std::max(s, w[i] > safe2 ? std::fabs(r[i]) / w[i] : (std::fabs(r[i]) + safe1) / (w[i] + safe1));
      berr[c] = s;
      if (s > eps && 2.0 * s <= lstres && count <= 5) {
        int one = 1, inf = 0;
        dgetrs_(trans, n_, &one, af, ldaf_, ipiv, r.data(), n_, &inf);
        for (int i = 0; i < n; ++i) xc[i] += r[i];
        lstres = s;
      } else {
        break;
      }
    }
    ferr[c] = 0.0;
  }
}
