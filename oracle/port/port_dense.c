/* port_dense.c -- the small dense linear algebra the reference takes from Eigen, restated in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see port.h).
 *   port_llt        Eigen::LLT<MatrixXd>::compute, info() == NumericalIssue when a pivot is <= 0
 *                   (used at Gradient_admm.h:38-41, Optimization3D_admm.h:313-316, Optimization3D_multi.h:703-706)
 *   port_llt_solve  LLT::solve
 *   port_min_eig    eigenvalues()(0) of Eigen::SelfAdjointEigenSolver (Gradient_admm.h:44-46): only the smallest
 *                   eigenvalue is consumed; cyclic Jacobi converges to it to round-off.
 * Eigen's blocked kernels sum in a different order than these loops: results agree to round-off (1e-12 relative in the
 * tests), the SPD / not-SPD decision is the same except within round-off of a singular matrix.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "port.h"

int port_llt(const double *A, double *L, int n) {
  memset(L, 0, sizeof(double) * (size_t)n * n);
  for (int k = 0; k < n; k++) {
    double x = A[k + (size_t)n * k];
    for (int j = 0; j < k; j++) x -= L[k + (size_t)n * j] * L[k + (size_t)n * j];
    if (!(x > 0)) return 0;
    double lkk = sqrt(x);
    L[k + (size_t)n * k] = lkk;
    for (int i = k + 1; i < n; i++) {
      double s = A[i + (size_t)n * k];
      for (int j = 0; j < k; j++) s -= L[i + (size_t)n * j] * L[k + (size_t)n * j];
      L[i + (size_t)n * k] = s / lkk;
    }
  }
  return 1;
}

void port_llt_solve(const double *L, int n, double *x) {
  for (int k = 0; k < n; k++) {
    double s = x[k];
    for (int j = 0; j < k; j++) s -= L[k + (size_t)n * j] * x[j];
    x[k] = s / L[k + (size_t)n * k];
  }
  for (int k = n - 1; k >= 0; k--) {
    double s = x[k];
    for (int j = k + 1; j < n; j++) s -= L[j + (size_t)n * k] * x[j];
    x[k] = s / L[k + (size_t)n * k];
  }
}

double port_min_eig(const double *A_in, int n) {
  double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
  memcpy(A, A_in, sizeof(double) * (size_t)n * n);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int p = 0; p < n; p++) {
      diag += A[p + (size_t)n * p] * A[p + (size_t)n * p];
      for (int q = p + 1; q < n; q++) off += A[p + (size_t)n * q] * A[p + (size_t)n * q];
    }
    if (off == 0 || off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[p + (size_t)n * q];
        if (apq == 0) continue;
        double app = A[p + (size_t)n * p], aqq = A[q + (size_t)n * q];
        double theta = (aqq - app) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        double cs = 1 / sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < n; k++) {
          double akp = A[k + (size_t)n * p], akq = A[k + (size_t)n * q];
          A[k + (size_t)n * p] = cs * akp - sn * akq;
          A[k + (size_t)n * q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = A[p + (size_t)n * k], aqk = A[q + (size_t)n * k];
          A[p + (size_t)n * k] = cs * apk - sn * aqk;
          A[q + (size_t)n * k] = sn * apk + cs * aqk;
        }
      }
  }
  double mn = A[0];
  for (int p = 1; p < n; p++) mn = fmin(mn, A[p + (size_t)n * p]);
  free(A);
  return mn;
}
