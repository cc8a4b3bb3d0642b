// optplane.cuh -- per-pair Newton refinement of separating planes: the persistent-plane mode ("optimal_plane": 1).
//
//   optimal_cd       Optimal_plane::optimal_cd       (HighOrderCCD/Optimal_plane.h:160-293) with barrier_grad (:118-158),
//                    barrier_energy (:93-116), current_c (:74-80), current_d (:82-91): Newton on the two tangent angles
//                    of c for a (sub-segment, obstacle point) plane; d stays tied to the point: d = -c.q - offset
//   self_optimal_cd  Optimal_plane::self_optimal_cd  (:620-773) with self_barrier_grad (:556-618) and
//                    self_barrier_energy (:518-554): Newton on (theta, phi, d) for an inter-robot plane
//
// Both keep the reference's control flow: the Hessian entry of phi is identically 0 and its gradient entry too (the
// reference adds 0 to grad(1), :147 / :583), the 2x2 system is regularised with 1e-2 I, a failed Cholesky triggers the
// shift by the smallest eigenvalue (+1e-8), steps longer than pi/2 are clamped, 0.8 Armijo ladder.  The reference's
// loops are unbounded; here they are capped (max_outer / max_ladder) far beyond what the oracle ever executes and a
// capped exit is reported to the caller.
// One pair per thread: everything lives in registers, sin/cos/log are the CUDA libm ones (<= 1-2 ulp from glibc), so
// results are tolerance-matched (not bitwise) against the reference.
#pragma once
#include <math.h>

#ifndef TOB_HD
#if defined(__CUDACC__)
#define TOB_HD __host__ __device__ __forceinline__
#else
#define TOB_HD inline
#endif
#endif

namespace tob {

#define TOB_OPT_MAX_OUTER 100000
#define TOB_OPT_MAX_LADDER 2000
#define TOB_PI 3.14159265358979323846

TOB_HD void op_tangent_frame(const double* c, double* c0, double* c1) {
  // c0 = (c.y, -c.x, 0) normalised; c1 = c0 x c normalised (Optimal_plane.h:175-179)
  double n0 = sqrt(c[1] * c[1] + c[0] * c[0] + 0.0);
  c0[0] = c[1] / n0; c0[1] = -c[0] / n0; c0[2] = 0.0 / n0;
  double x = c0[1] * c[2] - c0[2] * c[1], y = c0[2] * c[0] - c0[0] * c[2], z = c0[0] * c[1] - c0[1] * c[0];
  double n1 = sqrt(x * x + y * y + z * z);
  c1[0] = x / n1; c1[1] = y / n1; c1[2] = z / n1;
}

TOB_HD void op_current_c(const double* c, const double* c0, const double* c1, double theta, double phi, double* out) {
  const double ct = cos(theta), st = sin(theta), cp = cos(phi), sp = sin(phi);
  for (int k = 0; k < 3; ++k) out[k] = ct * c[k] + st * (cp * c0[k] + sp * c1[k]);
}

TOB_HD double op_barrier(double dist, double margin) { return -(dist - margin) * (dist - margin) * log(dist / margin); }

// Optimal_plane::barrier_energy: d is recomputed from c (current_d)
TOB_HD double op_energy_point(const double (*P)[3], const double* q, const double* c, double offset, double margin, double* d_out) {
  const double d = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - offset;
  *d_out = d;
  double e = 0;
  for (int j = 0; j < 6; ++j) {
    const double dist = (P[j][0] * c[0] + P[j][1] * c[1] + P[j][2] * c[2]) + d;
    if (dist <= 0) return INFINITY;
    if (dist < margin) e += op_barrier(dist, margin);
  }
  return e;
}

TOB_HD double op_energy_hulls(const double (*P0)[3], const double (*P1)[3], const double* c, double d, double offset, double margin) {
  double e = 0;
  for (int j = 0; j < 6; ++j) {
    const double dist = (P0[j][0] * c[0] + P0[j][1] * c[1] + P0[j][2] * c[2]) + d - 0.5 * offset;
    if (dist <= 0) return INFINITY;
    if (dist < margin) e += op_barrier(dist, margin);
  }
  for (int j = 0; j < 6; ++j) {
    const double dist = -(P1[j][0] * c[0] + P1[j][1] * c[1] + P1[j][2] * c[2]) - d - 0.5 * offset;
    if (dist <= 0) return INFINITY;
    if (dist < margin) e += op_barrier(dist, margin);
  }
  return e;
}

TOB_HD void op_e1e2(double dist, double margin, double* e1, double* e2) {
  const double lg = log(dist / margin), dm = dist - margin;
  *e1 = -(2 * dm * lg + dm * dm / dist);
  *e2 = -(2 * lg + 4 * dm / dist - dm * dm / (dist * dist));
}

// Smallest eigenvalue of a symmetric 2x2 / 3x3 matrix the way Eigen 3.3.7 computes `SelfAdjointEigenSolver(H).eigenvalues()(0)`
// (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h:395-443 compute(), :482-545 computeFromTridiagonal_impl, :806-864
// tridiagonal_qr_step; Tridiagonalization.h:456-497 for the 3x3 reduction; Jacobi/Jacobi.h:215-250 makeGivens): scale the
// lower triangle into [-1,1], reduce to tridiagonal form, implicit symmetric QR with the Wilkinson shift until the
// sub-diagonal deflates, scale back.  The caller shifts the Hessian by this value plus 1e-8, so the Newton direction is
// sensitive to its last bits: the operation order of Eigen is kept.
TOB_HD void op_givens(double p, double q, double* c, double* s) {
  if (q == 0.0) { *c = p < 0.0 ? -1.0 : 1.0; *s = 0.0; }
  else if (p == 0.0) { *c = 0.0; *s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) {
    const double t = q / p;
    double u = sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    *c = 1.0 / u; *s = -t * *c;
  } else {
    const double t = p / q;
    double u = sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    *s = -1.0 / u; *c = -t * *s;
  }
}

template <int N>
TOB_HD double op_min_eig_tridiag(double* diag, double* sub, double scale) {
  const double tiny = 2.2250738585072014e-308, prec = 2.0 * 2.220446049250313e-16;
  int end = N - 1, start = 0, iter = 0;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabs(sub[i]) <= (fabs(diag[i]) + fabs(diag[i + 1])) * prec || fabs(sub[i]) <= tiny) sub[i] = 0.0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    if (++iter > 30 * N) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    // one implicit QR step on [start, end]
    const double td = (diag[end - 1] - diag[end]) * 0.5, e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) mu -= fabs(e);
    else {
      const double e2 = e * e;
      const double ax = fabs(td), ay = fabs(e), pm = ax > ay ? ax : ay, qm = (ay < ax ? ay : ax) / pm;
      const double h = pm * sqrt(1.0 + qm * qm);                       // numext::hypot(td, e)
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu, z = sub[start];
    for (int k = start; k < end; ++k) {
      double c, s;
      op_givens(x, z, &c, &s);
      const double sdk = s * diag[k] + c * sub[k];
      const double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
    }
  }
  double m = diag[0];
  for (int i = 1; i < N; ++i) if (diag[i] < m) m = diag[i];
  return m * scale;
}

TOB_HD double op_min_eig2(double a, double b, double d) {
  double scale = fmax(fabs(a), fmax(fabs(b), fabs(d)));
  if (scale == 0.0) scale = 1.0;
  double diag[2] = {a / scale, d / scale}, sub[1] = {b / scale};
  return op_min_eig_tridiag<2>(diag, sub, scale);
}

TOB_HD double op_min_eig3(const double* H) {   // H: 3x3 symmetric, row-major; the lower triangle is used
  double scale = fmax(fmax(fabs(H[0]), fabs(H[3])), fmax(fmax(fabs(H[4]), fabs(H[6])), fmax(fabs(H[7]), fabs(H[8]))));
  if (scale == 0.0) scale = 1.0;
  const double m00 = H[0] / scale, m10 = H[3] / scale, m11 = H[4] / scale, m20 = H[6] / scale, m21 = H[7] / scale, m22 = H[8] / scale;
  double diag[3], sub[2];
  diag[0] = m00;
  const double v1norm2 = m20 * m20;
  if (v1norm2 <= 2.2250738585072014e-308) {
    diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
  } else {
    const double beta = sqrt(m10 * m10 + v1norm2), inv_beta = 1.0 / beta;
    const double m01 = m10 * inv_beta, m02 = m20 * inv_beta;
    const double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
  }
  return op_min_eig_tridiag<3>(diag, sub, scale);
}

// Eigen::LLT (unblocked, lower): fails on a pivot <= 0
TOB_HD bool op_llt2(double a, double b, double d, double* L) {
  if (a <= 0) return false;
  L[0] = sqrt(a); L[1] = b / L[0];
  const double x = d - L[1] * L[1];
  if (x <= 0) return false;
  L[2] = sqrt(x);
  return true;
}

TOB_HD bool op_llt3(const double* H, double* L) {   // L: l00 l10 l11 l20 l21 l22
  double x = H[0];
  if (x <= 0) return false;
  L[0] = sqrt(x); L[1] = H[3] / L[0]; L[3] = H[6] / L[0];
  x = H[4] - L[1] * L[1];
  if (x <= 0) return false;
  L[2] = sqrt(x); L[4] = (H[7] - L[3] * L[1]) / L[2];
  x = H[8] - (L[3] * L[3] + L[4] * L[4]);
  if (x <= 0) return false;
  L[5] = sqrt(x);
  return true;
}

// returns 0 = converged by one of the reference's two exits, 1 = a loop cap was hit
TOB_HD int optimal_cd(const double (*P)[3], const double* q, double offset, double margin, double* c, double* d_io) {
  double d = *d_io;
  for (int outer = 0; outer < TOB_OPT_MAX_OUTER; ++outer) {
    double c0[3], c1[3];
    op_tangent_frame(c, c0, c1);
    double g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0;
    for (int j = 0; j < 6; ++j) {
      const double r[3] = {P[j][0] + (-q[0]), P[j][1] + (-q[1]), P[j][2] + (-q[2])};
      const double p_c = r[0] * c[0] + r[1] * c[1] + r[2] * c[2];
      const double dist = p_c - offset;
      if (dist < margin) {
        const double p_c0 = r[0] * c0[0] + r[1] * c0[1] + r[2] * c0[2];
        const double p_c1 = r[0] * c1[0] + r[1] * c1[1] + r[2] * c1[2];
        double e1, e2;
        op_e1e2(dist, margin, &e1, &e2);
        g0 += e1 * p_c0;
        g1 += 0;
        h00 += e2 * p_c0 * p_c0 - e1 * p_c;
        h01 += e1 * p_c1;
        h11 += 0;
      }
    }
    if (sqrt(g0 * g0 + g1 * g1) < 1e-2) {
      d = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - offset;
      *d_io = d;
      return 0;
    }
    h00 += 1e-2; h11 += 1e-2;
    double L[3];
    if (!op_llt2(h00, h01, h11, L)) {
      const double ev = op_min_eig2(h00, h01, h11);
      if (ev < 0) { h00 = h00 - ev + 1e-8; h11 = h11 - ev + 1e-8; }
      if (!op_llt2(h00, h01, h11, L)) { *d_io = d; return 1; }   // the reference would use an invalid factor here
    }
    // direction = -LLT.solve(grad)
    const double y0 = g0 / L[0], y1 = (g1 - L[1] * y0) / L[2];
    const double x1 = y1 / L[2], x0 = (y0 - L[1] * x1) / L[0];
    const double dir0 = -x0, dir1 = -x1;
    const double w = -(g0 * dir0 + g1 * dir1);
    double step = 1.0;
    if (fabs(dir0) > 0.5 * TOB_PI || fabs(dir1) > 0.5 * TOB_PI)
      step = 0.95 * fmin(0.5 * fabs(TOB_PI / dir0), 0.5 * fabs(TOB_PI / dir1));
    double tc[3], td;
    op_current_c(c, c0, c1, 0.0, 0.0, tc);
    const double e0 = op_energy_point(P, q, tc, offset, margin, &td);
    op_current_c(c, c0, c1, 0.0 + step * dir0, 0.0 + step * dir1, tc);
    double e1 = op_energy_point(P, q, tc, offset, margin, &td);
    int ladder = 0;
    while (e0 - 1e-4 * w * step < e1) {
      if (++ladder > TOB_OPT_MAX_LADDER) { *d_io = d; return 1; }
      step *= 0.8;
      op_current_c(c, c0, c1, 0.0 + step * dir0, 0.0 + step * dir1, tc);
      e1 = op_energy_point(P, q, tc, offset, margin, &td);
    }
    c[0] = tc[0]; c[1] = tc[1]; c[2] = tc[2];
    d = -(c[0] * q[0] + c[1] * q[1] + c[2] * q[2]) - offset;
    if (fabs((e1 - e0) / e0) < 1e-1) { *d_io = d; return 0; }
  }
  *d_io = d;
  return 1;
}

TOB_HD int self_optimal_cd(const double (*P0)[3], const double (*P1)[3], double offset, double margin, double* c, double* d_io) {
  double d = *d_io;
  for (int outer = 0; outer < TOB_OPT_MAX_OUTER; ++outer) {
    double c0[3], c1[3];
    op_tangent_frame(c, c0, c1);
    double g[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int side = 0; side < 2; ++side) {
      const double (*Q)[3] = side ? P1 : P0;
      const double sg = side ? -1.0 : 1.0;
      for (int j = 0; j < 6; ++j) {
        const double pc = Q[j][0] * c[0] + Q[j][1] * c[1] + Q[j][2] * c[2];
        const double dist = side ? (-pc - d - 0.5 * offset) : (pc + d - 0.5 * offset);
        if (dist < margin) {
          const double p_c = sg * pc;
          const double p_c0 = sg * (Q[j][0] * c0[0] + Q[j][1] * c0[1] + Q[j][2] * c0[2]);
          const double p_c1 = sg * (Q[j][0] * c1[0] + Q[j][1] * c1[1] + Q[j][2] * c1[2]);
          double e1, e2;
          op_e1e2(dist, margin, &e1, &e2);
          g[0] += e1 * p_c0; g[1] += 0; g[2] += sg * e1;
          H[0] += e2 * p_c0 * p_c0 - e1 * p_c; H[1] += e1 * p_c1; H[2] += sg * e2 * p_c0;
          H[3] += e1 * p_c1;                                      H[6] += sg * e2 * p_c0;
          H[8] += e2;
        }
      }
    }
    if (sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) < 1e-2) { *d_io = d; return 0; }
    double L[6];
    if (!op_llt3(H, L)) {
      const double ev = op_min_eig3(H);
      if (ev < 0) { H[0] = H[0] - ev + 1e-8; H[4] = H[4] - ev + 1e-8; H[8] = H[8] - ev + 1e-8; }
      if (!op_llt3(H, L)) { *d_io = d; return 1; }
    }
    const double y0 = g[0] / L[0], y1 = (g[1] - L[1] * y0) / L[2], y2 = (g[2] - L[3] * y0 - L[4] * y1) / L[5];
    const double x2 = y2 / L[5], x1 = (y1 - L[4] * x2) / L[2], x0 = (y0 - L[1] * x1 - L[3] * x2) / L[0];
    const double dir[3] = {-x0, -x1, -x2};
    const double w = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
    double step = 1.0;
    if (fabs(dir[0]) > 0.5 * TOB_PI || fabs(dir[1]) > 0.5 * TOB_PI)
      step = 0.95 * fmin(0.5 * fabs(TOB_PI / dir[0]), 0.5 * fabs(TOB_PI / dir[1]));
    double tc[3], td = d;
    op_current_c(c, c0, c1, 0.0, 0.0, tc);
    const double e0 = op_energy_hulls(P0, P1, tc, td, offset, margin);
    op_current_c(c, c0, c1, 0.0 + step * dir[0], 0.0 + step * dir[1], tc);
    td = d + step * dir[2];
    double e1 = op_energy_hulls(P0, P1, tc, td, offset, margin);
    int ladder = 0;
    while (e0 - 1e-4 * w * step < e1) {
      if (++ladder > TOB_OPT_MAX_LADDER) { *d_io = d; return 1; }
      step *= 0.8;
      op_current_c(c, c0, c1, 0.0 + step * dir[0], 0.0 + step * dir[1], tc);
      td = d + step * dir[2];
      e1 = op_energy_hulls(P0, P1, tc, td, offset, margin);
    }
    c[0] = tc[0]; c[1] = tc[1]; c[2] = tc[2];
    d = td;
  }
  *d_io = d;
  return 1;
}

}  // namespace tob
