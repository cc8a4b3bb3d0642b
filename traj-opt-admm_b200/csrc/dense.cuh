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

// ---- register-resident warp versions (compile-time size N <= 32) ---------------------------------------------------
// Lane i keeps row i of the matrix in registers; values travel between lanes with shuffles, so there is no shared-memory
// round trip and no __syncwarp in the dependency chain.  The outer loops stay ROLLED (straight-line code that runs once
// is instruction-fetch bound): the register window is shifted by one column per step so that the pivot column is always
// w[0] and every register index is static.  Operands and operation order are those of the shared-memory versions above.

// Cholesky with Eigen::LLT's pivot rule.  w[j] = A(lane, j) on entry (destroyed).  If Lsm != nullptr the factor is stored
// there (col-major, ld N, lower triangle).  All 32 lanes must call; the return value is uniform.
template <int N>
__device__ __forceinline__ bool warp_chol_roll(double (&w)[N], double* Lsm) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
#pragma unroll 1
  for (int k = 0; k < N; k++) {
    const double x = __shfl_sync(full, w[0], k);
    if (!(x > 0)) return false;
    const double lkk = sqrt(x);
    double lik = w[0] / lkk;
    if (lane == k) lik = lkk;
    if (Lsm && lane >= k && lane < N) Lsm[lane + N * k] = lik;
#pragma unroll
    for (int jj = 1; jj < N; jj++) {
      const double ljk = __shfl_sync(full, lik, (k + jj) & 31);
      if (k + jj < N && lane >= k + jj) w[jj] -= lik * ljk;
    }
#pragma unroll
    for (int jj = 0; jj + 1 < N; jj++) w[jj] = w[jj + 1];
  }
  return true;
}

// L L^T x = b, L in shared memory (col-major, ld N) as left by warp_chol_roll.  Lane i passes b_i and receives x_i; the
// vector stays in registers, the divisions by the diagonal are one reciprocal per lane.
template <int N>
__device__ __forceinline__ double warp_chol_solve_sm(const double* Lsm, double b) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const double rd = lane < N ? 1.0 / Lsm[lane + N * lane] : 0.0;
  double s = b;
#pragma unroll 1
  for (int k = 0; k < N; k++) {
    const double yk = __shfl_sync(full, s * rd, k);
    if (lane == k) s = yk;
    else if (lane > k && lane < N) s -= Lsm[lane + N * k] * yk;
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; k--) {
    const double xk = __shfl_sync(full, s * rd, k);
    if (lane == k) s = xk;
    else if (lane < k) s -= Lsm[k + N * lane] * xk;
  }
  return s;
}

// smallest eigenvalue of the symmetric matrix whose row `lane` is in w[] (full rows; destroyed): Householder
// tridiagonalisation with shuffles, then Sturm-count multisection (32 shifts per round) on the division-free
// three-term recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (sign changes = eigenvalues below x), rescaled
// against overflow.  d, e: shared scratch of N doubles each.  All 32 lanes must call; the result is uniform.
template <int N>
__device__ __forceinline__ double warp_min_eig_roll(double (&w)[N], double* d, double* e) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
#pragma unroll 1
  for (int k = 0; k + 2 < N; k++) {
    const double xi = (lane > k && lane < N) ? w[0] : 0.0;
    double sigma = xi * xi;
#pragma unroll
    for (int o = 16; o; o >>= 1) sigma += __shfl_xor_sync(full, sigma, o);
    const double x0 = __shfl_sync(full, w[0], k + 1);
    const double akk = __shfl_sync(full, w[0], k);
    const double tail = sigma - x0 * x0;
    if (lane == 0) d[k] = akk;
    if (tail > 0) {                       // else: column already tridiagonal (uniform)
      const double alpha = (x0 >= 0 ? -1.0 : 1.0) * sqrt(sigma);
      double vi = xi;
      if (lane == k + 1) vi = x0 - alpha;
      const double vn2 = tail + (x0 - alpha) * (x0 - alpha);
      const double beta = 2.0 / vn2;
      double pi = 0;
#pragma unroll
      for (int jj = 1; jj < N; jj++) {
        const double vj = __shfl_sync(full, vi, (k + jj) & 31);
        if (k + jj < N) pi += w[jj] * vj;
      }
      if (!(lane > k && lane < N)) pi = 0;
      pi *= beta;
      double kk = pi * vi;
#pragma unroll
      for (int o = 16; o; o >>= 1) kk += __shfl_xor_sync(full, kk, o);
      kk *= 0.5 * beta;
      const double wi = pi - kk * vi;
#pragma unroll
      for (int jj = 1; jj < N; jj++) {
        const double wj = __shfl_sync(full, wi, (k + jj) & 31), vj = __shfl_sync(full, vi, (k + jj) & 31);
        if (k + jj < N) w[jj] -= vi * wj + wi * vj;
      }
      if (lane == 0) e[k] = alpha;
    } else if (lane == 0) e[k] = x0;
#pragma unroll
    for (int jj = 0; jj + 1 < N; jj++) w[jj] = w[jj + 1];
  }
  {   // window now starts at column N-2
    const double dn2 = __shfl_sync(full, w[0], N - 2), en2 = __shfl_sync(full, w[0], N - 1);
    const double dn1 = __shfl_sync(full, w[1], N - 1);
    if (lane == 0) { d[N - 2] = dn2; e[N - 2] = en2; d[N - 1] = dn1; e[N - 1] = 0.0; }
  }
  __syncwarp();
  // Gershgorin lower bound; lambda_min <= min d_i
  double lo = INFINITY, hi = INFINITY;
  if (lane < N) {
    const double r = fabs(e[lane]) + (lane > 0 ? fabs(e[lane - 1]) : 0.0);
    lo = d[lane] - r;
    hi = d[lane];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(full, lo, o));
    hi = fmin(hi, __shfl_xor_sync(full, hi, o));
  }
  if (!(hi > lo)) return hi;
#pragma unroll 1
  for (int round = 0; round < 14; round++) {
    const double x = lo + (hi - lo) * ((lane + 1) / 33.0);
    int cnt = 0;
    double p0 = 1.0, p1 = d[0] - x;
    bool neg = p1 < 0;                     // sign of the last non-zero term
    if (neg) cnt++;
#pragma unroll 1
    for (int i = 1; i < N; i++) {
      const double ei = e[i - 1];
      double p2 = (d[i] - x) * p1 - (ei * ei) * p0;
      const double m = fabs(p2);
      if (m > 1e150) { p2 *= 1e-150; p1 *= 1e-150; }
      else if (m < 1e-150 && m > 0) { p2 *= 1e150; p1 *= 1e150; }
      if (p2 != 0) {
        const bool n2 = p2 < 0;
        if (n2 != neg) cnt++;
        neg = n2;
      }
      p0 = p1; p1 = p2;
    }
    const unsigned mk = __ballot_sync(full, cnt >= 1);
    double nlo = lo, nhi = hi;
    if (mk == 0) nlo = __shfl_sync(full, x, 31);
    else {
      const int f = __ffs(mk) - 1;
      nhi = __shfl_sync(full, x, f);
      if (f > 0) nlo = __shfl_sync(full, x, f - 1);
    }
    lo = nlo; hi = nhi;
    if (!(hi > lo)) break;
  }
  return 0.5 * (lo + hi);
}

}  // namespace tob
