// tables.cu -- one-time host set-up of the constant tables the kernels read.
//
// The reference keeps this inside its executables and utility templates; the formulas AND the floating-point
// evaluation order are reproduced so that the subdivision matrices are bit-identical to the reference's
// `subdivide_tree` (they feed the bit-exact broadphase):
//   binomials            HighOrderCCD/Utils/CCDUtils.h:110-138   (Combination<40>)
//   C2 conversion        :140-170                                  (Conversion<5>::convert_matrix)
//   jerk Gram matrix     :172-226                                  (Dynamic3D<5,3>::dynamic_matrix, +1e-8 I)
//   blossom subdivision  :228-315                                  (Blossom<5>::coefficient)
//   basis = blossom(a,b) * convert[i]   Main/admmPathPlanning3D.cpp:303-313
//   k-DOP axes           HighOrderCCD/Utils/CCDUtils.cpp:56-119, normalised as in admmPathPlanning3D.cpp:403-410
#include <math.h>

#include "ctx.cuh"

namespace tob {

static const int kAxes[49][3] = {
    {1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 1},   {1, -1, 1}, {1, 1, -1}, {1, -1, -1}, {0, 1, 1},  {0, 1, -1}, {1, 0, 1},
    {1, 0, -1}, {1, 1, 0},  {1, -1, 0}, {0, 2, 1},   {0, 2, -1}, {0, 1, 2},  {0, 1, -2},  {2, 0, 1},  {2, 0, -1}, {1, 0, 2},
    {1, 0, -2}, {2, 1, 0},  {2, -1, 0}, {1, 2, 0},   {1, -2, 0}, {1, 2, 1},  {1, 2, -1},  {1, -2, 1}, {-1, 2, 1}, {1, 1, 2},
    {1, 1, -2}, {1, -1, 2}, {-1, 1, 2}, {2, 1, 1},   {2, 1, -1}, {2, -1, 1}, {-2, 1, 1},  {2, 2, 1},  {2, 2, -1}, {2, -2, 1},
    {-2, 2, 1}, {2, 1, 2},  {2, 1, -2}, {2, -1, 2},  {-2, 1, 2}, {1, 2, 2},  {1, 2, -2},  {1, -2, 2}, {-1, 2, 2}};

// col-major 6x6 helpers
static inline double& at(double* m, int r, int c) { return m[r + 6 * c]; }
static inline double at(const double* m, int r, int c) { return m[r + 6 * c]; }

// out = A * B with Eigen's coefficient-based product order: sum over k from 0 upward
static void mul66(const double* A, const double* B, double* out) {
  for (int c = 0; c < 6; c++)
    for (int r = 0; r < 6; r++) {
      double acc = 0;
      for (int k = 0; k < 6; k++) acc += at(A, r, k) * at(B, k, c);
      at(out, r, c) = acc;
    }
}

int make_tables_host(const tob_params& p, const double* time_weight, std::vector<double>& basis, std::vector<double>& weight,
                     std::vector<double>& convert, std::vector<double>& mdyn, std::vector<double>& kdop) {
  const int N = 5, K = 3;
  const int P = p.piece_num, R = p.res;
  // Pascal triangle, integer arithmetic like the reference (temp = temp*(i-j)/(j+1))
  long comb[41][41] = {{0}};
  comb[0][0] = 1;
  for (int i = 1; i <= 40; i++) {
    long long t = 1;
    for (int j = 0; j <= i; j++) { comb[i][j] = (long)t; t = t * (i - j) / (j + 1); }
  }
  // conversion matrices
  convert.assign((size_t)P * 36, 0.0);
  for (int i = 0; i < P; i++)
    for (int k = 0; k < 6; k++) at(&convert[36 * i], k, k) = 1.0;
  for (int i = 0; i < P - 1; i++) {
    double w0 = time_weight ? time_weight[i] : 1.0, w1 = time_weight ? time_weight[i + 1] : 1.0;
    double pp = w0 / (w0 + w1), qq = w1 / (w0 + w1);
    double I0[2][3] = {{qq * qq, 2 * pp * qq, pp * pp}, {0, qq, pp}};
    double I1[2][3] = {{qq, pp, 0}, {qq * qq, 2 * pp * qq, pp * pp}};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        at(&convert[36 * i], N - 1 + r, N - 2 + c) = I1[r][c];
        at(&convert[36 * (i + 1)], r, c) = I0[r][c];
      }
  }
  // jerk Gram matrix
  mdyn.assign(36, 0.0);
  for (int i = 0; i <= N; i++)
    for (int j = 0; j <= N; j++) {
      double acc = 0;
      for (int k0 = 0; k0 <= K; k0++)
        for (int k1 = 0; k1 <= K; k1++) {
          if (i - k0 <= N - K && j - k1 <= N - K && i - k0 >= 0 && j - k1 >= 0) {
            double t = ((k0 + k1) % 2 == 0) ? 1 : -1;
            t *= comb[K][k0] * comb[K][k1] * comb[N - K][i - k0] * comb[N - K][j - k1] / (double)comb[2 * N - K - K][i + j - k0 - k1];
            for (int s = 0; s < K; s++) t *= (N - s) * (N - s);
            t /= (double)(2 * N - K - K + 1);
            acc += t;
          }
        }
      at(mdyn.data(), i, j) = acc;
    }
  for (int k = 0; k < 6; k++) at(mdyn.data(), k, k) = at(mdyn.data(), k, k) + 1e-8 * 1.0;

  // blossom subdivision x conversion
  basis.assign((size_t)P * R * 36, 0.0);
  weight.assign((size_t)P * R, 0.0);
  for (int k = 0; k < R; k++) {
    double t0 = k / double(R), t1 = (k + 1) / double(R);
    double pt0[6], pt1[6], q0[6], q1[6];
    double a0 = 1, a1 = 1, b0 = 1, b1 = 1;
    for (int i = 0; i <= N; i++) {
      pt0[i] = a0; a0 *= t0;
      q0[i] = b0;  b0 *= 1 - t0;
      pt1[i] = a1; a1 *= t1;
      q1[i] = b1;  b1 *= 1 - t1;
    }
    double M[36];
    for (int i = 0; i < 36; i++) M[i] = 0;
    for (int i = 0; i <= N; i++)
      for (int j = 0; j <= N; j++) {
        double acc = 0;
        if (i + j < N) {
          int mk = i < j ? i : j;
          for (int m = 0; m <= mk; m++)
            acc += comb[N - i][j - m] * comb[i][m] * q0[N - i - j + m] * q1[i - m] * pt0[j - m] * pt1[m];
        } else {
          int mk = (N - i) < (N - j) ? (N - i) : (N - j);
          for (int m = 0; m <= mk; m++)
            acc += comb[N - i][m] * comb[i][N - j - m] * q0[m] * q1[N - j - m] * pt0[N - i - m] * pt1[i + j - N + m];
        }
        at(M, i, j) = acc;
      }
    for (int i = 0; i < P; i++) {
      mul66(M, &convert[36 * i], &basis[(size_t)36 * (i * R + k)]);
      weight[i * R + k] = t1 - t0;
    }
  }
  // k-DOP axes, normalised (v /= sqrt(x*x+y*y+z*z))
  kdop.assign(3 * 49, 0.0);
  for (int k = 0; k < 49; k++) {
    double x = kAxes[k][0], y = kAxes[k][1], z = kAxes[k][2];
    double nrm = sqrt(x * x + y * y + z * z);
    kdop[3 * k] = x / nrm; kdop[3 * k + 1] = y / nrm; kdop[3 * k + 2] = z / nrm;
  }
  return 0;
}

}  // namespace tob
