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

// ---- warp-cooperative versions (one full warp, matrix in shared memory, n <= 32) --------------------------------

// same pivot rule as chol_is_spd_n, right-looking: lane i owns row i.  A (col-major, ld n) is overwritten with L in
// its lower triangle.  All 32 lanes must call; returns the same value on every lane.
__device__ inline bool warp_chol_is_spd(double* A, int n) {
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < n; k++) {
    double x = A[k + n * k];
    if (!(x > 0)) return false;           // uniform: every lane reads the same shared value
    double lkk = sqrt(x);
    double lik = 0;
    if (lane > k && lane < n) lik = A[lane + n * k] / lkk;
    __syncwarp();
    if (lane > k && lane < n) A[lane + n * k] = lik;
    if (lane == k) A[k + n * k] = lkk;
    __syncwarp();
    if (lane > k && lane < n)
      for (int j = k + 1; j <= lane; j++) A[lane + n * j] -= lik * A[j + n * k];
    __syncwarp();
  }
  return true;
}

// smallest eigenvalue of the symmetric n x n matrix A (full storage, col-major, ld n; destroyed):
// Householder tridiagonalisation by the warp, then Sturm-sequence multisection (32 shifts per round).
// d, e, v, w: shared scratch of n doubles each.  All 32 lanes must call; result identical on every lane.
__device__ inline double warp_min_eig(double* A, int n, double* d, double* e, double* v, double* w) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  for (int k = 0; k + 2 < n; k++) {
    double xi = (lane > k && lane < n) ? A[lane + n * k] : 0.0;
    double sigma = xi * xi;
    for (int o = 16; o; o >>= 1) sigma += __shfl_xor_sync(full, sigma, o);
    double x0 = A[k + 1 + n * k];
    double tail = sigma - x0 * x0;
    if (lane == 0) d[k] = A[k + n * k];
    if (!(tail > 0)) {                    // column already tridiagonal
      if (lane == 0) e[k] = x0;
      __syncwarp();
      continue;
    }
    double alpha = (x0 >= 0 ? -1.0 : 1.0) * sqrt(sigma);
    double vi = xi;
    if (lane == k + 1) vi = x0 - alpha;
    double vn2 = tail + (x0 - alpha) * (x0 - alpha);
    double beta = 2.0 / vn2;
    if (lane < n) v[lane] = (lane > k) ? vi : 0.0;
    __syncwarp();
    double pi = 0;
    if (lane > k && lane < n) {
      for (int j = k + 1; j < n; j++) pi += A[lane + n * j] * v[j];
      pi *= beta;
    }
    double kk = pi * vi;
    for (int o = 16; o; o >>= 1) kk += __shfl_xor_sync(full, kk, o);
    kk *= 0.5 * beta;
    double wi = pi - kk * vi;
    if (lane < n) w[lane] = (lane > k) ? wi : 0.0;
    __syncwarp();
    if (lane > k && lane < n)
      for (int j = k + 1; j < n; j++) A[lane + n * j] -= vi * w[j] + wi * v[j];
    if (lane == 0) e[k] = alpha;
    __syncwarp();
  }
  if (lane == 0) {
    if (n >= 2) { d[n - 2] = A[(n - 2) + n * (n - 2)]; e[n - 2] = A[(n - 1) + n * (n - 2)]; }
    d[n - 1] = A[(n - 1) + n * (n - 1)];
    e[n - 1] = 0.0;
  }
  __syncwarp();
  // Gershgorin lower bound; lambda_min <= min d_i
  double lo = INFINITY, hi = INFINITY;
  if (lane < n) {
    double r = fabs(e[lane]) + (lane > 0 ? fabs(e[lane - 1]) : 0.0);
    lo = d[lane] - r;
    hi = d[lane];
  }
  for (int o = 16; o; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(full, lo, o));
    hi = fmin(hi, __shfl_xor_sync(full, hi, o));
  }
  if (!(hi > lo)) return hi;
  for (int round = 0; round < 14; round++) {
    double x = lo + (hi - lo) * ((lane + 1) / 33.0);
    // number of eigenvalues < x
    int cnt = 0;
    double q = d[0] - x;
    if (q < 0) cnt++;
    for (int i = 1; i < n; i++) {
      if (q == 0) q = 1e-300;
      q = d[i] - x - e[i - 1] * e[i - 1] / q;
      if (q < 0) cnt++;
    }
    unsigned m = __ballot_sync(full, cnt >= 1);
    double nlo = lo, nhi = hi;
    if (m == 0) nlo = __shfl_sync(full, x, 31);
    else {
      int f = __ffs(m) - 1;
      nhi = __shfl_sync(full, x, f);
      if (f > 0) nlo = __shfl_sync(full, x, f - 1);
    }
    lo = nlo; hi = nhi;
    if (!(hi > lo)) break;
  }
  return 0.5 * (lo + hi);
}

}  // namespace tob
