// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's host+device math header (csrc/gjk.cuh) and the host-side table set-up (csrc/tables.cu)
// with g++ so that their bit-exactness against the compiled reference can be checked on a machine without a GPU.
// Nothing in the product links this file; the product path always runs the CUDA build of the same header.
#include <cstring>
#include <vector>

#include "../../traj-opt-admm_b200/csrc/gjk.cuh"
#include "../../traj-opt-admm_b200/csrc/ctx.cuh"
#include "../../traj-opt-admm_b200/csrc/optplane.cuh"

using namespace tob;

template <int N>
static void load(const double* src, double (*dst)[3]) {  // N x 3 column-major
  for (int j = 0; j < N; j++)
    for (int k = 0; k < 3; k++) dst[j][k] = src[k * N + j];
}

extern "C" {

int hs_gjk(const double* A, int na, const double* B, int nb, double* v) {
  if (na == 6 && nb == 1) { double a[6][3], b[1][3]; load<6>(A, a); load<1>(B, b); gjk_witness<6, 1>(a, b, v); return 0; }
  if (na == 12 && nb == 1) { double a[12][3], b[1][3]; load<12>(A, a); load<1>(B, b); gjk_witness<12, 1>(a, b, v); return 0; }
  if (na == 6 && nb == 6) { double a[6][3], b[6][3]; load<6>(A, a); load<6>(B, b); gjk_witness<6, 6>(a, b, v); return 0; }
  if (na == 12 && nb == 12) { double a[12][3], b[12][3]; load<12>(A, a); load<12>(B, b); gjk_witness<12, 12>(a, b, v); return 0; }
  return 1;
}

int hs_kdop_dcd(const double* P, const double* q, const double* kdop, double d) {
  double a[6][3], b[1][3] = {{q[0], q[1], q[2]}};
  load<6>(P, a);
  double lo[49], hi[49];
  kdop_extents<6>(a, kdop, lo, hi);
  bool r1 = kdop_point_overlap(lo, hi, kdop, q, d);
  bool r2 = kdop_overlap<6, 1>(a, b, kdop, d);
  return (r1 ? 1 : 0) | (r2 ? 2 : 0);   // both formulations must agree: 0 or 3
}

int hs_self_kdop_dcd(const double* P0, const double* P1, const double* kdop, double d) {
  double a[6][3], b[6][3];
  load<6>(P0, a); load<6>(P1, b);
  double lo0[49], hi0[49], lo1[49], hi1[49];
  kdop_extents<6>(a, kdop, lo0, hi0);
  kdop_extents<6>(b, kdop, lo1, hi1);
  bool r1 = kdop_sets_overlap(lo0, hi0, lo1, hi1, d);
  bool r2 = kdop_overlap<6, 6>(a, b, kdop, d);
  return (r1 ? 1 : 0) | (r2 ? 2 : 0);
}

int hs_kdop_ccd(const double* P, const double* D, const double* q, const double* kdop, double d, double t0, double t1) {
  double p[6][3], dd[6][3], A[12][3], b[1][3] = {{q[0], q[1], q[2]}};
  load<6>(P, p); load<6>(D, dd);
  swept_points(p, dd, t0, t1, A);
  return kdop_overlap<12, 1>(A, b, kdop, d);
}

int hs_gjk_ccd(const double* P, const double* D, const double* q, double d, double t0, double t1) {
  double p[6][3], dd[6][3], A[12][3], b[1][3] = {{q[0], q[1], q[2]}}, v[3];
  load<6>(P, p); load<6>(D, dd);
  swept_points(p, dd, t0, t1, A);
  gjk_witness<12, 1>(A, b, v);
  return (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) <= d * d;
}

int hs_self_gjk_ccd(const double* P0, const double* D0, const double* P1, const double* D1, double d, double s0, double s1) {
  double p0[6][3], d0[6][3], p1[6][3], d1[6][3], A[12][3], B[12][3], v[3];
  load<6>(P0, p0); load<6>(D0, d0); load<6>(P1, p1); load<6>(D1, d1);
  swept_points(p0, d0, 0.0, s0, A);
  swept_points(p1, d1, 0.0, s1, B);
  gjk_witness<12, 12>(A, B, v);
  return (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) <= d * d;
}

int hs_plane_point(const double* P, const double* q, double dist, double offset, double* c, double* d) {
  double a[6][3];
  load<6>(P, a);
  return plane_point(a, q, dist, offset, c, d);
}

int hs_plane_hulls(const double* P0, const double* P1, double dist, double* c, double* d) {
  double a[6][3], b[6][3];
  load<6>(P0, a); load<6>(P1, b);
  return plane_hulls(a, b, dist, c, d);
}

int hs_refine_d(const double* P0, const double* P1, const double* c, double offset, double margin, double* d) {
  double a[6][3], b[6][3];
  load<6>(P0, a); load<6>(P1, b);
  return refine_d(a, b, c, offset, margin, d, 10000);
}

int hs_optimal_cd(const double* P, const double* q, double offset, double margin, double* c, double* d) {
  double a[6][3];
  load<6>(P, a);
  return optimal_cd(a, q, offset, margin, c, d);
}

int hs_self_optimal_cd(const double* P0, const double* P1, double offset, double margin, double* c, double* d) {
  double a[6][3], b[6][3];
  load<6>(P0, a); load<6>(P1, b);
  return self_optimal_cd(a, b, offset, margin, c, d);
}

int hs_make_tables(int piece_num, int res, double* basis, double* weight, double* convert, double* mdyn, double* kdop) {
  tob_params p;
  std::memset(&p, 0, sizeof(p));
  p.piece_num = piece_num; p.res = res; p.uav_num = 1;
  std::vector<double> b, w, cv, md, kd;
  make_tables_host(p, nullptr, b, w, cv, md, kd);
  std::memcpy(basis, b.data(), b.size() * sizeof(double));
  std::memcpy(weight, w.data(), w.size() * sizeof(double));
  std::memcpy(convert, cv.data(), cv.size() * sizeof(double));
  std::memcpy(mdyn, md.data(), 36 * sizeof(double));
  std::memcpy(kdop, kd.data(), 147 * sizeof(double));
  return 0;
}

// single-precision filter of the 49-DOP gate against the plain FP64 gate: thresholds made exactly like segments.cu (k_rows)
// for one row P (6 x 3 column-major), then n points.  out_f = filtered decision (stage 0..14 && stage 14..49 like k_narrow),
// out_x = kdop_point_overlap; *n_exact = axes the filter handed to the FP64 fallback.
int hs_kdop_gate_batch(const double* P, const double* pts, int n, const double* kdop, double d, unsigned char* out_f,
                       unsigned char* out_x, unsigned long long* n_exact) {
  double a[6][3];
  load<6>(P, a);
  double lo[TOB_KDOP_AXES], hi[TOB_KDOP_AXES];
  kdop_extents<6>(a, kdop, lo, hi);
  double m[3];
  for (int k = 0; k < 3; k++) {
    double l = INFINITY, h = -INFINITY;
    for (int j = 0; j < 6; j++) { if (a[j][k] < l) l = a[j][k]; if (a[j][k] > h) h = a[j][k]; }
    m[k] = 0.5 * (l + h);
  }
  float kf[TOB_KF_ROW], kdf[3 * TOB_KDOP_AXES], tmag = 0.f;
  for (int k = 0; k < 3 * TOB_KDOP_AXES; k++) kdf[k] = (float)kdop[k];
  for (int k = 0; k < TOB_KDOP_AXES; k++) {
    const double x = kdop[3 * k], y = kdop[3 * k + 1], z = kdop[3 * k + 2];
    const float tm = kdop_gate_thresholds(lo[k], hi[k], d, x * m[0] + y * m[1] + z * m[2], &kf[2 * k], &kf[2 * k + 1]);
    tmag = fmaxf(tmag, tm);
    if (!(tm <= 3.0e38f)) tmag = INFINITY;
  }
  kf[2 * TOB_KDOP_AXES] = kdop_gate_allowance(tmag, m, d);
  kf[2 * TOB_KDOP_AXES + 1] = 0.f;
  unsigned groups = 0, exact = 0;
  for (int i = 0; i < n; i++) {
    const double q[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    out_f[i] = kdop_point_gate(kf, m, kdf, lo, hi, kdop, q, d, &groups, &exact, 0, 14) &&
               kdop_point_gate(kf, m, kdf, lo, hi, kdop, q, d, &groups, &exact, 14, TOB_KDOP_AXES);
    out_x[i] = kdop_point_overlap(lo, hi, kdop, q, d);
  }
  *n_exact = exact;
  return 0;
}

}  // extern "C"

// ---- fastlog.cuh: the barrier kernels' logarithm, host build (tests/test_cpu_checks.py::test_fast_log) ----
#include "../../traj-opt-admm_b200/csrc/fastlog.cuh"
extern "C" void hs_fast_log(const double* x, int n, double* out) {
  for (int i = 0; i < n; i++) out[i] = tob::tob_log_pos(x[i], tob::kLogTabHost);
}
extern "C" void hs_ref_logl(const double* x, int n, double* out_hi, double* out_lo) {
  // long-double logarithm split in two doubles (64-bit mantissa: 11 more bits than the value under test)
  for (int i = 0; i < n; i++) {
    long double l = logl((long double)x[i]);
    out_hi[i] = (double)l;
    out_lo[i] = (double)(l - (long double)out_hi[i]);
  }
}

// ---- the narrowphase decision of k_narrow in the product's own arithmetic, both orders, for one row and n points
// (tests/test_cpu_checks.py::test_hostsim_narrow_rule): ref = all 49 axes, then GJK (the reference's order);
// cut = the first `gate1` axes, GJK, the remaining axes only when the witness is not shorter than skip = dist (1 - 1e-6)
extern "C" int hs_narrow_rule(const double* P, const double* pts, int n, const double* kdop, double dist, double offset, int gate1,
                              unsigned char* ok_ref, double* pl_ref, unsigned char* ok_cut, double* pl_cut, unsigned long long* n_band,
                              unsigned long long* groups_ref, unsigned long long* groups_cut) {
  double a[6][3];
  load<6>(P, a);
  double lo[TOB_KDOP_AXES], hi[TOB_KDOP_AXES];
  kdop_extents<6>(a, kdop, lo, hi);
  const double skip = dist * (1.0 - 1e-6);
  unsigned gr = 0, gc = 0;
  unsigned long long band = 0;
  for (int i = 0; i < n; i++) {
    const double q[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    double c[3], d;
    ok_ref[i] = 0; ok_cut[i] = 0;
    for (int k = 0; k < 4; k++) { pl_ref[4 * i + k] = 0; pl_cut[4 * i + k] = 0; }
    if (kdop_point_overlap(lo, hi, kdop, q, dist, &gr) && plane_point(a, q, dist, offset, c, &d)) {
      ok_ref[i] = 1; pl_ref[4 * i] = c[0]; pl_ref[4 * i + 1] = c[1]; pl_ref[4 * i + 2] = c[2]; pl_ref[4 * i + 3] = d;
    }
    if (kdop_point_overlap(lo, hi, kdop, q, dist, &gc, 0, gate1)) {
      const double cn = plane_point_witness(a, q, c);
      bool acc = !(cn > dist);
      if (acc && !(cn <= skip)) { band++; acc = kdop_point_overlap(lo, hi, kdop, q, dist, &gc, gate1, TOB_KDOP_AXES); }
      if (acc) {
        plane_point_finish(q, offset, cn, c, &d);
        ok_cut[i] = 1; pl_cut[4 * i] = c[0]; pl_cut[4 * i + 1] = c[1]; pl_cut[4 * i + 2] = c[2]; pl_cut[4 * i + 3] = d;
      }
    }
  }
  *n_band = band; *groups_ref = gr; *groups_cut = gc;
  return 0;
}
