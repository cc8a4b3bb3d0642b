/* port_optplane.c -- plain-C restatement of the persistent-plane refinement ("optimal_plane": 1).
 *
 * TEST INFRASTRUCTURE ONLY (see port.h).
 *
 *   port_optimal_cd_impl       Optimal_plane::optimal_cd      (HighOrderCCD/Optimal_plane.h:160-293)
 *                              with barrier_grad :118-158, barrier_energy :93-116, current_c :74-80, current_d :82-91
 *   port_self_optimal_cd_impl  Optimal_plane::self_optimal_cd (:620-773) with self_barrier_grad :556-618,
 *                              self_barrier_energy :518-554
 *   small_min_eig              Eigen 3.3.7 SelfAdjointEigenSolver<Matrix2d/Matrix3d>(H).eigenvalues()(0):
 *                              lib/eigen3/Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h:395-443 (scaling), :482-545
 *                              (deflation loop), :806-864 (implicit QR step with the Wilkinson shift),
 *                              Tridiagonalization.h:456-497 (3x3 reduction), Jacobi/Jacobi.h:215-250 (Givens rotation)
 *   llt2 / llt3                Eigen::LLT, unblocked lower factorisation; "NumericalIssue" = a pivot <= 0
 *
 * The reference loops are `while(true)`; the port keeps them unbounded except for a very large guard that only protects
 * the test-suite from a hang (returns 1 when hit). */
#include <math.h>
#include <float.h>

#include "port.h"

#define G g_port
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define GUARD_OUTER 100000000L

static void tangent_frame(const double *c, double *c0, double *c1) {        /* :175-179 */
  double n0 = sqrt(c[1] * c[1] + c[0] * c[0] + 0.0);
  c0[0] = c[1] / n0; c0[1] = -c[0] / n0; c0[2] = 0.0 / n0;
  double x = c0[1] * c[2] - c0[2] * c[1], y = c0[2] * c[0] - c0[0] * c[2], z = c0[0] * c[1] - c0[1] * c[0];
  double n1 = sqrt(x * x + y * y + z * z);
  c1[0] = x / n1; c1[1] = y / n1; c1[2] = z / n1;
}
static void current_c(const double *c, const double *c0, const double *c1, double theta, double phi, double *out) { /* :74-80 */
  for (int k = 0; k < 3; k++) out[k] = cos(theta) * c[k] + sin(theta) * (cos(phi) * c0[k] + sin(phi) * c1[k]);
}
static double bar(double dist) { return -(dist - G.margin) * (dist - G.margin) * log(dist / G.margin); }
static void e1e2(double dist, double *e1, double *e2) {
  const double m = G.margin;
  *e1 = -(2 * (dist - m) * log(dist / m) + (dist - m) * (dist - m) / dist);
  *e2 = -(2 * log(dist / m) + 4 * (dist - m) / dist - (dist - m) * (dist - m) / (dist * dist));
}
static double energy_point(const double (*P)[3], const double *q, const double *c) {   /* :93-116 */
  double d = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - G.offset, e = 0;
  for (int j = 0; j < 6; j++) {
    double dist = (P[j][0] * c[0] + P[j][1] * c[1] + P[j][2] * c[2]) + d;
    if (dist <= 0) return INFINITY;
    if (dist < G.margin) e += bar(dist);
  }
  return e;
}
static double energy_hulls(const double (*P0)[3], const double (*P1)[3], const double *c, double d) {   /* :518-554 */
  double e = 0;
  for (int j = 0; j < 6; j++) {
    double dist = (P0[j][0] * c[0] + P0[j][1] * c[1] + P0[j][2] * c[2]) + d - 0.5 * G.offset;
    if (dist <= 0) return INFINITY;
    if (dist < G.margin) e += bar(dist);
  }
  for (int j = 0; j < 6; j++) {
    double dist = -(P1[j][0] * c[0] + P1[j][1] * c[1] + P1[j][2] * c[2]) - d - 0.5 * G.offset;
    if (dist <= 0) return INFINITY;
    if (dist < G.margin) e += bar(dist);
  }
  return e;
}

/* ---- Eigen's small symmetric eigenvalue path ---------------------------------------------------------------------------- */
static void givens(double p, double q, double *c, double *s) {              /* Jacobi.h:215-250 */
  if (q == 0.0) { *c = p < 0.0 ? -1.0 : 1.0; *s = 0.0; }
  else if (p == 0.0) { *c = 0.0; *s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) { double t = q / p, u = sqrt(1.0 + t * t); if (p < 0.0) u = -u; *c = 1.0 / u; *s = -t * *c; }
  else { double t = p / q, u = sqrt(1.0 + t * t); if (q < 0.0) u = -u; *s = -1.0 / u; *c = -t * *s; }
}
static double eig_hypot(double x, double y) {                               /* MathFunctionsImpl.h:73-83 */
  x = fabs(x); y = fabs(y);
  double p = x > y ? x : y;
  if (p == 0.0) return 0.0;
  double qp = (y < x ? y : x) / p;
  return p * sqrt(1.0 + qp * qp);
}
static double tridiag_min(double *diag, double *sub, int n, double scale) { /* SelfAdjointEigenSolver.h:482-545, :806-864 */
  const double prec = 2.0 * DBL_EPSILON;
  int end = n - 1, start = 0, iter = 0;
  while (end > 0) {
    for (int i = start; i < end; i++)
      if (fabs(sub[i]) <= (fabs(diag[i]) + fabs(diag[i + 1])) * prec || fabs(sub[i]) <= DBL_MIN) sub[i] = 0.0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    double td = (diag[end - 1] - diag[end]) * 0.5, e = sub[end - 1], mu = diag[end];
    if (td == 0.0) mu -= fabs(e);
    else {
      double e2 = e * e, h = eig_hypot(td, e);
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu, z = sub[start];
    for (int k = start; k < end; k++) {
      double c, s;
      givens(x, z, &c, &s);
      double sdk = s * diag[k] + c * sub[k], dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
    }
  }
  double m = diag[0];
  for (int i = 1; i < n; i++) if (diag[i] < m) m = diag[i];
  return m * scale;
}
static double min_eig2(double a, double b, double d) {
  double scale = fmax(fabs(a), fmax(fabs(b), fabs(d)));
  if (scale == 0.0) scale = 1.0;
  double diag[2] = {a / scale, d / scale}, sub[1] = {b / scale};
  return tridiag_min(diag, sub, 2, scale);
}
static double min_eig3(const double *H) {   /* row-major symmetric 3x3; lower triangle */
  double scale = 0;
  const int low[6] = {0, 3, 4, 6, 7, 8};
  for (int i = 0; i < 6; i++) if (fabs(H[low[i]]) > scale) scale = fabs(H[low[i]]);
  if (scale == 0.0) scale = 1.0;
  double m00 = H[0] / scale, m10 = H[3] / scale, m11 = H[4] / scale, m20 = H[6] / scale, m21 = H[7] / scale, m22 = H[8] / scale;
  double diag[3], sub[2];
  diag[0] = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= DBL_MIN) { diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21; }   /* Tridiagonalization.h:466-474 */
  else {
    double beta = sqrt(m10 * m10 + v1norm2), inv = 1.0 / beta, m01 = m10 * inv, m02 = m20 * inv;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q; diag[2] = m22 - m02 * q; sub[0] = beta; sub[1] = m21 - m01 * q;
  }
  return tridiag_min(diag, sub, 3, scale);
}
static int llt2(double a, double b, double d, double *L) {
  if (a <= 0) return 0;
  L[0] = sqrt(a); L[1] = b / L[0];
  double x = d - L[1] * L[1];
  if (x <= 0) return 0;
  L[2] = sqrt(x);
  return 1;
}
static int llt3(const double *H, double *L) {
  double x = H[0];
  if (x <= 0) return 0;
  L[0] = sqrt(x); L[1] = H[3] / L[0]; L[3] = H[6] / L[0];
  x = H[4] - L[1] * L[1];
  if (x <= 0) return 0;
  L[2] = sqrt(x); L[4] = (H[7] - L[3] * L[1]) / L[2];
  x = H[8] - (L[3] * L[3] + L[4] * L[4]);
  if (x <= 0) return 0;
  L[5] = sqrt(x);
  return 1;
}

/* ---- Optimal_plane::optimal_cd (:160-293) --------------------------------------------------------------------------------- */
int port_optimal_cd_impl(const double (*P)[3], const double *q, double *c, double *d_io) {
  const double offset = G.offset, margin = G.margin;
  double d = *d_io;
  for (long outer = 0; outer < GUARD_OUTER; outer++) {
    double c0[3], c1[3];
    tangent_frame(c, c0, c1);
    double g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0;
    for (int j = 0; j < 6; j++) {                                             /* barrier_grad :118-158 */
      double r[3] = {P[j][0] + (-q[0]), P[j][1] + (-q[1]), P[j][2] + (-q[2])};
      double p_c = r[0] * c[0] + r[1] * c[1] + r[2] * c[2], dist = p_c - offset;
      if (dist < margin) {
        double p_c0 = r[0] * c0[0] + r[1] * c0[1] + r[2] * c0[2], p_c1 = r[0] * c1[0] + r[1] * c1[1] + r[2] * c1[2], e1, e2;
        e1e2(dist, &e1, &e2);
        g0 += e1 * p_c0; g1 += 0;
        h00 += e2 * p_c0 * p_c0 - e1 * p_c; h01 += e1 * p_c1; h11 += 0;
      }
    }
    if (sqrt(g0 * g0 + g1 * g1) < 1e-2) { *d_io = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - offset; return 0; }   /* :187-193 */
    h00 += 1e-2; h11 += 1e-2;                                                /* :197-199 */
    double L[3];
    if (!llt2(h00, h01, h11, L)) {                                            /* :201-212 */
      double ev = min_eig2(h00, h01, h11);
      if (ev < 0) { h00 = h00 - ev + 1e-8; h11 = h11 - ev + 1e-8; }
      if (!llt2(h00, h01, h11, L)) { *d_io = d; return 1; }
    }
    double y0 = g0 / L[0], y1 = (g1 - L[1] * y0) / L[2], x1 = y1 / L[2], x0 = (y0 - L[1] * x1) / L[0];
    double dir0 = -x0, dir1 = -x1, w = -(g0 * dir0 + g1 * dir1), step = 1.0;
    if (fabs(dir0) > 0.5 * M_PI || fabs(dir1) > 0.5 * M_PI) step = 0.95 * fmin(0.5 * fabs(M_PI / dir0), 0.5 * fabs(M_PI / dir1));
    double tc[3];
    current_c(c, c0, c1, 0.0, 0.0, tc);
    double e0 = energy_point(P, q, tc);
    current_c(c, c0, c1, 0.0 + step * dir0, 0.0 + step * dir1, tc);
    double e1 = energy_point(P, q, tc);
    while (e0 - 1e-4 * w * step < e1) {                                       /* :246-258 */
      step *= 0.8;
      current_c(c, c0, c1, 0.0 + step * dir0, 0.0 + step * dir1, tc);
      e1 = energy_point(P, q, tc);
    }
    c[0] = tc[0]; c[1] = tc[1]; c[2] = tc[2];
    d = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - offset;
    if (fabs((e1 - e0) / e0) < 1e-1) { *d_io = d; return 0; }                  /* :286-287 */
  }
  *d_io = d;
  return 1;
}

/* ---- Optimal_plane::self_optimal_cd (:620-773) ---------------------------------------------------------------------------- */
int port_self_optimal_cd_impl(const double (*P0)[3], const double (*P1)[3], double *c, double *d_io) {
  const double offset = G.offset, margin = G.margin;
  double d = *d_io;
  for (long outer = 0; outer < GUARD_OUTER; outer++) {
    double c0[3], c1[3];
    tangent_frame(c, c0, c1);
    double g[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 6; j++) {                                             /* :568-591 */
      double pc = P0[j][0] * c[0] + P0[j][1] * c[1] + P0[j][2] * c[2], dist = pc + d - 0.5 * offset;
      if (dist < margin) {
        double p_c0 = P0[j][0] * c0[0] + P0[j][1] * c0[1] + P0[j][2] * c0[2], p_c1 = P0[j][0] * c1[0] + P0[j][1] * c1[1] + P0[j][2] * c1[2], e1, e2;
        e1e2(dist, &e1, &e2);
        g[0] += e1 * p_c0; g[1] += 0; g[2] += e1;
        H[0] += e2 * p_c0 * p_c0 - e1 * pc; H[1] += e1 * p_c1; H[2] += e2 * p_c0; H[3] += e1 * p_c1; H[6] += e2 * p_c0; H[8] += e2;
      }
    }
    for (int j = 0; j < 6; j++) {                                             /* :592-616 */
      double pc = -(P1[j][0] * c[0] + P1[j][1] * c[1] + P1[j][2] * c[2]), dist = pc - d - 0.5 * offset;
      if (dist < margin) {
        double p_c0 = -(P1[j][0] * c0[0] + P1[j][1] * c0[1] + P1[j][2] * c0[2]), p_c1 = -(P1[j][0] * c1[0] + P1[j][1] * c1[1] + P1[j][2] * c1[2]), e1, e2;
        e1e2(dist, &e1, &e2);
        g[0] += e1 * p_c0; g[1] += 0; g[2] += -e1;
        H[0] += e2 * p_c0 * p_c0 - e1 * pc; H[1] += e1 * p_c1; H[2] += -e2 * p_c0; H[3] += e1 * p_c1; H[6] += -e2 * p_c0; H[8] += e2;
      }
    }
    if (sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) < 1e-2) { *d_io = d; return 0; }       /* :641-642 */
    double L[6];
    if (!llt3(H, L)) {                                                        /* :644-660 */
      double ev = min_eig3(H);
      if (ev < 0) { H[0] = H[0] - ev + 1e-8; H[4] = H[4] - ev + 1e-8; H[8] = H[8] - ev + 1e-8; }
      if (!llt3(H, L)) { *d_io = d; return 1; }
    }
    double y0 = g[0] / L[0], y1 = (g[1] - L[1] * y0) / L[2], y2 = (g[2] - L[3] * y0 - L[4] * y1) / L[5];
    double x2 = y2 / L[5], x1 = (y1 - L[4] * x2) / L[2], x0 = (y0 - L[1] * x1 - L[3] * x2) / L[0];
    double dir[3] = {-x0, -x1, -x2}, w = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]), step = 1.0;
    if (fabs(dir[0]) > 0.5 * M_PI || fabs(dir[1]) > 0.5 * M_PI) step = 0.95 * fmin(0.5 * fabs(M_PI / dir[0]), 0.5 * fabs(M_PI / dir[1]));
    double tc[3], td = d;
    current_c(c, c0, c1, 0.0, 0.0, tc);
    double e0 = energy_hulls(P0, P1, tc, td);
    current_c(c, c0, c1, 0.0 + step * dir[0], 0.0 + step * dir[1], tc);
    td = d + step * dir[2];
    double e1 = energy_hulls(P0, P1, tc, td);
    while (e0 - 1e-4 * w * step < e1) {                                       /* :742-754 */
      step *= 0.8;
      current_c(c, c0, c1, 0.0 + step * dir[0], 0.0 + step * dir[1], tc);
      td = d + step * dir[2];
      e1 = energy_hulls(P0, P1, tc, td);
    }
    c[0] = tc[0]; c[1] = tc[1]; c[2] = tc[2];
    d = td;
  }
  *d_io = d;
  return 1;
}

static void load6c(const double *src, double (*P)[3]) {
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) P[j][k] = src[k * 6 + j];
}
void port_optimal_cd(const double *P, const double *q, double *c, double *d) {
  double A[6][3];
  load6c(P, A);
  port_optimal_cd_impl(A, q, c, d);
}
void port_self_optimal_cd(const double *P0, const double *P1, double *c, double *d) {
  double A[6][3], B[6][3];
  load6c(P0, A); load6c(P1, B);
  port_self_optimal_cd_impl(A, B, c, d);
}
