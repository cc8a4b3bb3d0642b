// dense.cuh -- small dense FP64 helpers run by a single thread (19x19 / 13x13 blocks).
#pragma once
#include <math.h>

namespace tob {

// Eigen::LLT pivot rule (Eigen/src/Cholesky/LLT.h, unblocked kernel): fails when a_kk - sum l_kj^2 <= 0.
// A, L column-major with leading dimension n.
__device__ inline bool chol_is_spd_n(const double* A, double* L, int n) {
  for (int k = 0; k < n; k++) {
    double x = A[k + n * k];
    for (int j = 0; j < k; j++) x -= L[k + n * j] * L[k + n * j];
    if (!(x > 0)) return false;
    double lkk = sqrt(x);
    L[k + n * k] = lkk;
    for (int i = k + 1; i < n; i++) {
      double s = A[i + n * k];
      for (int j = 0; j < k; j++) s -= L[i + n * j] * L[k + n * j];
      L[i + n * k] = s / lkk;
    }
  }
  return true;
}

// smallest eigenvalue of a symmetric n x n matrix by cyclic Jacobi; A is destroyed.
// (the reference uses Eigen::SelfAdjointEigenSolver; only lambda_min is consumed, Gradient_admm.h:44-52)
__device__ inline double jacobi_min_eig_n(double* A, int n) {
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int p = 0; p < n; p++) {
      diag += A[p + n * p] * A[p + n * p];
      for (int q = p + 1; q < n; q++) off += A[p + n * q] * A[p + n * q];
    }
    if (off == 0 || off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[p + n * q];
        if (apq == 0) continue;
        double app = A[p + n * p], aqq = A[q + n * q];
        double theta = (aqq - app) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        double cs = 1 / sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < n; k++) {
          double akp = A[k + n * p], akq = A[k + n * q];
          A[k + n * p] = cs * akp - sn * akq;
          A[k + n * q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = A[p + n * k], aqk = A[q + n * k];
          A[p + n * k] = cs * apk - sn * aqk;
          A[q + n * k] = sn * apk + cs * aqk;
        }
      }
  }
  double mn = A[0];
  for (int p = 1; p < n; p++) mn = fmin(mn, A[p + n * p]);
  return mn;
}

}  // namespace tob
