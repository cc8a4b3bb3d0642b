// segments.cu -- per-row geometry: control points of every Bezier sub-segment, its (swept) AABB and its 49-DOP
// extents.  One warp per row (robot x sub-segment).
//
// Reference arithmetic reproduced (built with --fmad=false):
//   P_tr = basis_tr * bz, bz = spline.block<6,3>(3*sp_id,0): Eigen coefficient-based product, sum over k from 0
//   upward with separate multiply and add (BVH.cpp:160-163, Optimization3D_admm.h:92-97, Step.h:62-70).
//   CCD: D_tr = basis_tr * bz_d  (Step.h:69-70);  swept box = AABB(P_tr U basis_tr*(bz+bz_d))  (BVH.cpp:209-243)
//   trial point of the line search: spline + step*direction element-wise (Optimization3D_admm.h:537).
//   k-DOP extents: level = x*Px + y*Py + z*Pz, left to right (CCD.h:376-378).
#include "ctx.cuh"
#include "gjk.cuh"

namespace tob {

struct RowArgs {
  const double* spline;   // robots x 3T
  const double* dir;      // robots x 3T or null
  const double* step;     // robots or null
  const double* basis;    // n_tr x 36
  const double* kdop;     // 147
  double *P, *D, *box, *klo, *khi;
  float* kf;              // rows x TOB_KF_ROW: single-precision filter of the 49-DOP gate (gjk.cuh: kdop_point_gate)
  double* kc;             // rows x 4: the centre the filter's thresholds refer to
  double gate_d;          // gap of the gate the thresholds are made for (offset + margin)
  int n_tr, res, T, row_begin, row_end, mode;
  int* kmax;              // mode & 2: per-robot CCD ladder exponent, reset here for the fused CCD kernel
};

__global__ void __launch_bounds__(128) k_rows(RowArgs a) {
  __shared__ double sP[4][18], sQ[4][18], sBz[4][18], sBd[4][18], sM[4][3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = a.row_begin + blockIdx.x * 4 + w;
  const bool live = row < a.row_end;
  const int robot = live ? row / a.n_tr : 0, tr = live ? row - robot * a.n_tr : 0;
  const int piece = tr / a.res;
  const double* B = a.basis + (size_t)36 * tr;
  if (live && (a.mode & 2) && a.kmax && tr == 0 && lane == 0) a.kmax[robot] = 0;
  if (live && lane < 18) {
    int m = lane % 6, ax = lane / 6;
    size_t g = (size_t)robot * 3 * a.T + (size_t)ax * a.T + 3 * piece + m;
    double s = a.spline[g];
    double dd = a.dir ? a.dir[g] : 0.0;
    if (a.mode & 4) { s = s + a.step[robot] * dd; }
    sBz[w][lane] = s;
    sBd[w][lane] = dd;
  }
  __syncwarp();
  if (live && lane < 18) {
    int j = lane % 6, ax = lane / 6;
    double acc = 0;
    for (int k = 0; k < 6; k++) acc += B[j + 6 * k] * sBz[w][k + 6 * ax];
    sP[w][lane] = acc;
    a.P[(size_t)18 * row + lane] = acc;
    if (a.mode & 2) {
      double ad = 0, aq = 0;
      for (int k = 0; k < 6; k++) ad += B[j + 6 * k] * sBd[w][k + 6 * ax];
      for (int k = 0; k < 6; k++) aq += B[j + 6 * k] * (sBz[w][k + 6 * ax] + sBd[w][k + 6 * ax]);
      a.D[(size_t)18 * row + lane] = ad;
      sQ[w][lane] = aq;
    }
  }
  __syncwarp();
  if (live && lane < 3) {
    double lo = INFINITY, hi = -INFINITY;
    for (int j = 0; j < 6; j++) {
      double v = sP[w][j + 6 * lane];
      if (v < lo) lo = v;
      if (v > hi) hi = v;
      if (a.mode & 2) {
        v = sQ[w][j + 6 * lane];
        if (v < lo) lo = v;
        if (v > hi) hi = v;
      }
    }
    a.box[(size_t)6 * row + lane] = lo;
    a.box[(size_t)6 * row + 3 + lane] = hi;
    sM[w][lane] = 0.5 * (lo + hi);      // centre of the filter's thresholds: any point near the row would do
  }
  __syncwarp();
  if (live && (a.mode & 1)) {
    float tmag = 0.f;
    for (int k = lane; k < TOB_KDOP_AXES; k += 32) {
      double x = a.kdop[3 * k], y = a.kdop[3 * k + 1], z = a.kdop[3 * k + 2];
      double lo = INFINITY, hi = -INFINITY;
      for (int j = 0; j < 6; j++) {
        double lv = x * sP[w][j] + y * sP[w][j + 6] + z * sP[w][j + 12];
        if (lv < lo) lo = lv;
        if (lv > hi) hi = lv;
      }
      a.klo[(size_t)TOB_KDOP_AXES * row + k] = lo;
      a.khi[(size_t)TOB_KDOP_AXES * row + k] = hi;
      float tl, th;
      const float tm = kdop_gate_thresholds(lo, hi, a.gate_d, x * sM[w][0] + y * sM[w][1] + z * sM[w][2], &tl, &th);
      tmag = fmaxf(tmag, tm);
      if (!(tm <= 3.0e38f)) tmag = INFINITY;     // NaN / overflow: the allowance becomes infinite, every axis is re-tested
      a.kf[(size_t)TOB_KF_ROW * row + 2 * k] = tl;
      a.kf[(size_t)TOB_KF_ROW * row + 2 * k + 1] = th;
    }
    for (int o = 16; o; o >>= 1) tmag = fmaxf(tmag, __shfl_xor_sync(0xffffffffu, tmag, o));
    if (lane == 0) {
      a.kf[(size_t)TOB_KF_ROW * row + 2 * TOB_KDOP_AXES] = kdop_gate_allowance(tmag, sM[w], a.gate_d);
      a.kf[(size_t)TOB_KF_ROW * row + 2 * TOB_KDOP_AXES + 1] = 0.f;
    }
    if (lane < 4) a.kc[(size_t)4 * row + lane] = lane < 3 ? sM[w][lane] : 0.0;
  }
}

int compute_rows(tob_ctx* c, const double* spline_dev, const double* dir_dev, const double* step_dev, int rb, int re, int mode) {
  int rows = c->rows_all();
  TOB_CUDA(c, c->geo.P.ensure((size_t)18 * rows));
  TOB_CUDA(c, c->geo.D.ensure((size_t)18 * rows));
  TOB_CUDA(c, c->geo.box.ensure((size_t)6 * rows));
  TOB_CUDA(c, c->geo.klo.ensure((size_t)TOB_KDOP_AXES * rows));
  TOB_CUDA(c, c->geo.khi.ensure((size_t)TOB_KDOP_AXES * rows));
  TOB_CUDA(c, c->geo.kf.ensure((size_t)TOB_KF_ROW * rows));
  TOB_CUDA(c, c->geo.kc.ensure((size_t)4 * rows));
  RowArgs a;
  a.spline = spline_dev; a.dir = dir_dev; a.step = step_dev;
  a.basis = c->d_basis.p; a.kdop = c->d_kdop.p;
  a.P = c->geo.P.p; a.D = c->geo.D.p; a.box = c->geo.box.p; a.klo = c->geo.klo.p; a.khi = c->geo.khi.p;
  a.kf = c->geo.kf.p; a.kc = c->geo.kc.p; a.gate_d = c->prm.offset + c->prm.margin;
  a.n_tr = c->n_tr; a.res = c->prm.res; a.T = c->T; a.row_begin = rb * c->n_tr; a.row_end = re * c->n_tr; a.mode = mode;
  a.kmax = c->kmax.p;
  if (re > rb) {
    Prof prof(c, K_ROWS);
    k_rows<<<div_up((re - rb) * c->n_tr, 4), 128, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

}  // namespace tob
