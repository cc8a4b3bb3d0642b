/* port.c -- plain-C restatement of traj-opt-admm's per-ADMM-iteration hot path (the CPU oracle, "port" flavour).
 *
 * TEST INFRASTRUCTURE ONLY (see port.h): the checker of the CUDA product, never a product path.
 * Exported symbols mirror oracle/ref_shim.cpp (prefix port_ instead of ref_) so that the parity tests can be run against
 * either oracle.  Every function cites the reference file:line it follows.  Known, documented differences from the
 * compiled reference: candidate lists are produced in ascending point-id order (the reference returns its aabb::Tree DFS
 * order; the contract is the SET, BVH/src/AABB.cc:131-161), hence floating-point sums over planes are taken in another
 * order (round-off only); inter-robot pairs are visited in lexicographic order; dense factorizations are unblocked.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "port.h"

port_ctx g_port;
#define G g_port

/* ---- set-up: Main/admmPathPlanning3D.cpp:249-353,403-414 and HighOrderCCD/Utils/CCDUtils.h:110-315 ------------------- */
static const int kAxes[KDOP_AXES][3] = {   /* HighOrderCCD/Utils/CCDUtils.cpp:56-119 */
    {1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 1},   {1, -1, 1}, {1, 1, -1}, {1, -1, -1}, {0, 1, 1},  {0, 1, -1}, {1, 0, 1},
    {1, 0, -1}, {1, 1, 0},  {1, -1, 0}, {0, 2, 1},   {0, 2, -1}, {0, 1, 2},  {0, 1, -2},  {2, 0, 1},  {2, 0, -1}, {1, 0, 2},
    {1, 0, -2}, {2, 1, 0},  {2, -1, 0}, {1, 2, 0},   {1, -2, 0}, {1, 2, 1},  {1, 2, -1},  {1, -2, 1}, {-1, 2, 1}, {1, 1, 2},
    {1, 1, -2}, {1, -1, 2}, {-1, 1, 2}, {2, 1, 1},   {2, 1, -1}, {2, -1, 1}, {-2, 1, 1},  {2, 2, 1},  {2, 2, -1}, {2, -2, 1},
    {-2, 2, 1}, {2, 1, 2},  {2, 1, -2}, {2, -1, 2},  {-2, 1, 2}, {1, 2, 2},  {1, 2, -2},  {1, -2, 2}, {-1, 2, 2}};

#define M6(m, r, c) ((m)[(r) + 6 * (c)])

void port_setup(int piece_num, int res, int uav_num, double lambda, double margin, double offset, double mu, double vel_limit,
                double acc_limit, double ks, double kt, int optimal_plane) {
  G.optimal_plane = optimal_plane;
  free(G.is_sep); free(G.sep_cd); free(G.is_self_sep); free(G.self_sep_cd);      /* sized by n_tr / uav_num: start empty */
  G.is_sep = NULL; G.sep_cd = NULL; G.is_self_sep = NULL; G.self_sep_cd = NULL;
  const int N = ORDER, K = 3;
  G.piece_num = piece_num; G.res = res; G.uav_num = uav_num; G.n_tr = piece_num * res;
  G.T = (N + 1) + (piece_num - 1) * (N + 1 - 3);                       /* admmPathPlanning3D.cpp:255 */
  G.lambda = lambda; G.margin = margin; G.offset = offset; G.mu = mu;
  G.vel_limit = vel_limit; G.acc_limit = acc_limit; G.ks = ks; G.kt = kt;
  G.wolfe = 0; G.gnorm = 1;
  free(G.basis); free(G.weight); free(G.convert);
  G.basis = (double *)calloc((size_t)G.n_tr * 36, sizeof(double));
  G.weight = (double *)calloc((size_t)G.n_tr, sizeof(double));
  G.convert = (double *)calloc((size_t)piece_num * 36, sizeof(double));
  /* Combination<40>::value (CCDUtils.h:110-138), integer arithmetic */
  static long comb[41][41];
  memset(comb, 0, sizeof(comb));
  comb[0][0] = 1;
  for (int i = 1; i <= 40; i++) {
    long long t = 1;
    for (int j = 0; j <= i; j++) { comb[i][j] = (long)t; t = t * (i - j) / (j + 1); }
  }
  /* Conversion<5>::convert_matrix (:140-170) with time_weight == 1 */
  for (int i = 0; i < piece_num; i++)
    for (int k = 0; k < 6; k++) M6(G.convert + 36 * i, k, k) = 1.0;
  for (int i = 0; i < piece_num - 1; i++) {
    double w0 = 1.0, w1 = 1.0;
    double p = w0 / (w0 + w1), q = w1 / (w0 + w1);
    double I0[2][3] = {{q * q, 2 * p * q, p * p}, {0, q, p}};
    double I1[2][3] = {{q, p, 0}, {q * q, 2 * p * q, p * p}};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        M6(G.convert + 36 * i, N - 1 + r, N - 2 + c) = I1[r][c];
        M6(G.convert + 36 * (i + 1), r, c) = I0[r][c];
      }
  }
  /* Dynamic3D<5,3>::dynamic_matrix (:172-226): Gram matrix of the third derivative + 1e-8 I */
  for (int i = 0; i <= N; i++)
    for (int j = 0; j <= N; j++) {
      double acc = 0;
      for (int k0 = 0; k0 <= K; k0++)
        for (int k1 = 0; k1 <= K; k1++)
          if (i - k0 <= N - K && j - k1 <= N - K && i - k0 >= 0 && j - k1 >= 0) {
            double t = ((k0 + k1) % 2 == 0) ? 1 : -1;
            t *= comb[K][k0] * comb[K][k1] * comb[N - K][i - k0] * comb[N - K][j - k1] / (double)comb[2 * N - K - K][i + j - k0 - k1];
            for (int s = 0; s < K; s++) t *= (N - s) * (N - s);
            t /= (double)(2 * N - K - K + 1);
            acc += t;
          }
      M6(G.mdyn, i, j) = acc;
    }
  for (int k = 0; k < 6; k++) M6(G.mdyn, k, k) = M6(G.mdyn, k, k) + 1e-8 * 1.0;
  /* Blossom<5>::coefficient (:228-315) times convert_list[i] (admmPathPlanning3D.cpp:303-313) */
  for (int k = 0; k < res; k++) {
    double t0 = k / (double)res, t1 = (k + 1) / (double)res;
    double pt0[6], pt1[6], q0[6], q1[6];
    double a0 = 1, a1 = 1, b0 = 1, b1 = 1;
    for (int i = 0; i <= N; i++) {
      pt0[i] = a0; a0 *= t0;
      q0[i] = b0;  b0 *= 1 - t0;
      pt1[i] = a1; a1 *= t1;
      q1[i] = b1;  b1 *= 1 - t1;
    }
    double Mb[36];
    for (int i = 0; i <= N; i++)
      for (int j = 0; j <= N; j++) {
        double acc = 0;
        if (i + j < N) {
          int mk = i < j ? i : j;
          for (int m = 0; m <= mk; m++) acc += comb[N - i][j - m] * comb[i][m] * q0[N - i - j + m] * q1[i - m] * pt0[j - m] * pt1[m];
        } else {
          int mk = (N - i) < (N - j) ? (N - i) : (N - j);
          for (int m = 0; m <= mk; m++)
            acc += comb[N - i][m] * comb[i][N - j - m] * q0[m] * q1[N - j - m] * pt0[N - i - m] * pt1[i + j - N + m];
        }
        M6(Mb, i, j) = acc;
      }
    for (int i = 0; i < piece_num; i++) {
      double *out = G.basis + (size_t)36 * (i * res + k);
      const double *Cv = G.convert + 36 * i;
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) {
          double acc = 0;
          for (int kk = 0; kk < 6; kk++) acc += M6(Mb, r, kk) * M6(Cv, kk, c);
          M6(out, r, c) = acc;
        }
      G.weight[i * res + k] = t1 - t0;
    }
  }
  /* normalised k-DOP axes (admmPathPlanning3D.cpp:403-410) */
  for (int k = 0; k < KDOP_AXES; k++) {
    double x = kAxes[k][0], y = kAxes[k][1], z = kAxes[k][2];
    double nrm = sqrt(x * x + y * y + z * z);
    G.kdop[3 * k] = x / nrm; G.kdop[3 * k + 1] = y / nrm; G.kdop[3 * k + 2] = z / nrm;
  }
}

int port_trajectory_num(void) { return G.T; }

void port_get_tables(double *basis, double *weight, double *convert, double *mdyn, double *kdop) {
  memcpy(basis, G.basis, sizeof(double) * 36 * (size_t)G.n_tr);
  memcpy(weight, G.weight, sizeof(double) * (size_t)G.n_tr);
  memcpy(convert, G.convert, sizeof(double) * 36 * (size_t)G.piece_num);
  memcpy(mdyn, G.mdyn, sizeof(double) * 36);
  memcpy(kdop, G.kdop, sizeof(double) * 3 * KDOP_AXES);
}

/* ---- cloud: BVH::InitPointcloud (HighOrderCCD/BVH/BVH.cpp:53-92).  The port keeps the points in a uniform grid; only
 *      the leaf predicate of the reference's tree is contractual. ------------------------------------------------------- */
static struct {
  double lo[3], h;
  int n[3];
  int *cell_start;   /* cells + 1 */
  int *cell_pts;     /* point ids, ascending inside a cell */
} grid;

void port_init_pointcloud(const double *V, int n) {
  free(G.is_sep); free(G.sep_cd); free(G.is_self_sep); free(G.self_sep_cd);      /* sized by n_pts: start empty */
  G.is_sep = NULL; G.sep_cd = NULL; G.is_self_sep = NULL; G.self_sep_cd = NULL;
  free(G.V);
  G.V = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  memcpy(G.V, V, sizeof(double) * 3 * (size_t)n);
  G.n_pts = n;
  double hi[3];
  for (int a = 0; a < 3; a++) {
    grid.lo[a] = INFINITY; hi[a] = -INFINITY;
    for (int i = 0; i < n; i++) {
      double v = V[(size_t)a * n + i];
      if (v < grid.lo[a]) grid.lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  }
  grid.h = 0.5;
  for (;;) {
    double cells = 1;
    for (int a = 0; a < 3; a++) { grid.n[a] = (int)floor((hi[a] - grid.lo[a]) / grid.h) + 1; cells *= grid.n[a]; }
    if (cells <= 4e6) break;
    grid.h *= 2;
  }
  size_t nc = (size_t)grid.n[0] * grid.n[1] * grid.n[2];
  free(grid.cell_start); free(grid.cell_pts);
  grid.cell_start = (int *)calloc(nc + 1, sizeof(int));
  grid.cell_pts = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *cell_of = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    int c[3];
    for (int a = 0; a < 3; a++) {
      c[a] = (int)floor((V[(size_t)a * n + i] - grid.lo[a]) / grid.h);
      if (c[a] < 0) c[a] = 0;
      if (c[a] >= grid.n[a]) c[a] = grid.n[a] - 1;
    }
    cell_of[i] = (c[2] * grid.n[1] + c[1]) * grid.n[0] + c[0];
    grid.cell_start[cell_of[i] + 1]++;
  }
  for (size_t k = 0; k < nc; k++) grid.cell_start[k + 1] += grid.cell_start[k];
  int *fill = (int *)malloc(sizeof(int) * nc);
  for (size_t k = 0; k < nc; k++) fill[k] = grid.cell_start[k];
  for (int i = 0; i < n; i++) grid.cell_pts[fill[cell_of[i]]++] = i;
  free(fill); free(cell_of);
}

static int cmp_uint(const void *a, const void *b) {
  unsigned x = *(const unsigned *)a, y = *(const unsigned *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* all points p with, for every axis, !(p + d < lo) && !(p > hi + d): aabb::AABB::overlaps(box, true, d)
 * (BVH/src/AABB.cc:131-161 as called at :647).  ids ascending.  Returns the count (ids may be NULL / too small). */
static long box_query(const double *lo, const double *hi, double d, unsigned *ids, long cap) {
  long n = 0;
  int c0[3], c1[3];
  for (int a = 0; a < 3; a++) {
    c0[a] = (int)floor((lo[a] - d - grid.lo[a]) / grid.h) - 1;
    c1[a] = (int)floor((hi[a] + d - grid.lo[a]) / grid.h) + 1;
    if (c0[a] < 0) c0[a] = 0;
    if (c1[a] >= grid.n[a]) c1[a] = grid.n[a] - 1;
  }
  const int np = G.n_pts;
  for (int z = c0[2]; z <= c1[2]; z++)
    for (int y = c0[1]; y <= c1[1]; y++)
      for (int x = c0[0]; x <= c1[0]; x++) {
        size_t cell = ((size_t)z * grid.n[1] + y) * grid.n[0] + x;
        for (int k = grid.cell_start[cell]; k < grid.cell_start[cell + 1]; k++) {
          int i = grid.cell_pts[k];
          int ok = 1;
          for (int a = 0; a < 3; a++) {
            double p = G.V[(size_t)a * np + i];
            if ((p + d < lo[a]) || (p > hi[a] + d)) { ok = 0; break; }
          }
          if (ok) { if (ids && n < cap) ids[n] = (unsigned)i; n++; }
        }
      }
  if (ids && n <= cap) qsort(ids, (size_t)n, sizeof(unsigned), cmp_uint);
  return n;
}

/* ---- sub-segment control points: P = basis_tr * bz, Eigen coefficient-based product, sum over k from 0
 *      (BVH.cpp:160-163, Optimization3D_admm.h:92-97).  P[j][axis]. ----------------------------------------------------- */
static void seg_points(const double *spline, int tr, double (*P)[3]) {
  const int sp = tr / G.res, T = G.T;
  const double *B = G.basis + (size_t)36 * tr;
  for (int ax = 0; ax < 3; ax++)
    for (int j = 0; j < 6; j++) {
      double acc = 0;
      for (int k = 0; k < 6; k++) acc += M6(B, j, k) * spline[(size_t)ax * T + 3 * sp + k];
      P[j][ax] = acc;
    }
}
/* basis_tr * (bz + bz_d) (BVH.cpp:209-211) */
static void seg_points_moved(const double *spline, const double *dir, int tr, double (*Q)[3]) {
  const int sp = tr / G.res, T = G.T;
  const double *B = G.basis + (size_t)36 * tr;
  for (int ax = 0; ax < 3; ax++)
    for (int j = 0; j < 6; j++) {
      double acc = 0;
      for (int k = 0; k < 6; k++) acc += M6(B, j, k) * (spline[(size_t)ax * T + 3 * sp + k] + dir[(size_t)ax * T + 3 * sp + k]);
      Q[j][ax] = acc;
    }
}

void port_segment_points(const double *spline, int tr_id, double *Pout) {
  double P[6][3];
  seg_points(spline, tr_id, P);
  for (int j = 0; j < 6; j++) for (int a = 0; a < 3; a++) Pout[j + 6 * a] = P[j][a];
}

static void box_of(const double (*P)[3], int n, double *lo, double *hi, int init) {
  for (int a = 0; a < 3; a++) {
    if (init) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    for (int j = 0; j < n; j++) {
      if (P[j][a] < lo[a]) lo[a] = P[j][a];
      if (P[j][a] > hi[a]) hi[a] = P[j][a];
    }
  }
}

/* ---- broadphase: BVH::DCDCollision (BVH.cpp:149-192), BVH::CCDCollision (:195-249) ----------------------------------- */
typedef struct { unsigned *off; unsigned *ids; long n, cap; } CandList;

static void cand_push_row(CandList *cl, int row, const double *lo, const double *hi, double d) {
  long room = cl->cap - cl->n;
  long k = box_query(lo, hi, d, (cl->ids && room > 0) ? cl->ids + cl->n : NULL, room > 0 ? room : 0);
  cl->n += k;
  cl->off[row + 1] = (unsigned)cl->n;
}

long port_dcd_collision(const double *spline, double d, unsigned *off, unsigned *ids, long cap) {
  CandList cl = {off, ids, 0, cap};
  off[0] = 0;
  for (int tr = 0; tr < G.n_tr; tr++) {
    double P[6][3], lo[3], hi[3];
    seg_points(spline, tr, P);
    box_of(P, 6, lo, hi, 1);
    cand_push_row(&cl, tr, lo, hi, d);
  }
  return cl.n;
}

long port_ccd_collision(const double *spline, const double *direction, double d, unsigned *off, unsigned *ids, long cap) {
  CandList cl = {off, ids, 0, cap};
  off[0] = 0;
  for (int tr = 0; tr < G.n_tr; tr++) {
    double P[6][3], Q[6][3], lo[3], hi[3];
    seg_points(spline, tr, P);
    seg_points_moved(spline, direction, tr, Q);
    box_of(P, 6, lo, hi, 1);
    box_of(Q, 6, lo, hi, 0);
    cand_push_row(&cl, tr, lo, hi, d);
  }
  return cl.n;
}

/* BVH::SelfDCDCollision (BVH.cpp:252-286) / SelfCCDCollision (:289-329): all pairs p0 < p1 whose (swept) boxes are within d;
 * leaf test of aabb::Tree::query(margin) (AABB.cc:691-698).  Pin/Din: u blocks of 6x3 col-major. */
static void load6(const double *src, double (*P)[3]) {
  for (int j = 0; j < 6; j++) for (int a = 0; a < 3; a++) P[j][a] = src[j + 6 * a];
}
static int boxes_within(const double *lo0, const double *hi0, const double *lo1, const double *hi1, double d) {
  for (int a = 0; a < 3; a++)
    if (hi0[a] + d < lo1[a] || lo0[a] > hi1[a] + d) return 0;
  return 1;
}
static long self_pairs(const double *Pin, const double *Din, int u, double d, unsigned *pairs, long cap) {
  double *lo = (double *)malloc(sizeof(double) * 3 * (size_t)u), *hi = (double *)malloc(sizeof(double) * 3 * (size_t)u);
  for (int i = 0; i < u; i++) {
    double P[6][3];
    load6(Pin + 18 * i, P);
    box_of(P, 6, lo + 3 * i, hi + 3 * i, 1);
    if (Din) {
      double D[6][3], Q[6][3];
      load6(Din + 18 * i, D);
      for (int j = 0; j < 6; j++) for (int a = 0; a < 3; a++) Q[j][a] = P[j][a] + D[j][a];
      box_of(Q, 6, lo + 3 * i, hi + 3 * i, 0);
    }
  }
  long n = 0;
  for (int a = 0; a < u; a++)
    for (int b = a + 1; b < u; b++)
      if (boxes_within(lo + 3 * a, hi + 3 * a, lo + 3 * b, hi + 3 * b, d)) {
        if (pairs && n < cap) { pairs[2 * n] = (unsigned)a; pairs[2 * n + 1] = (unsigned)b; }
        n++;
      }
  free(lo); free(hi);
  return n;
}
long port_self_dcd(const double *P, int u, double d, unsigned *pairs, long cap) { return self_pairs(P, NULL, u, d, pairs, cap); }
long port_self_ccd(const double *P, const double *D, int u, double d, unsigned *pairs, long cap) { return self_pairs(P, D, u, d, pairs, cap); }

/* ---- k-DOP tests: CCD::KDOPDCD (HighOrderCCD/CCD/CCD.h:354-413), SelfKDOPDCD (:535-587), KDOPCCD (:416-473),
 *      SelfKDOPCCD (:475-533).  level = x*px + y*py + z*pz, left to right (:376-378). ---------------------------------- */
static int kdop_overlap(const double (*A)[3], int na, const double (*B)[3], int nb, double d) {
  for (int k = 0; k < KDOP_AXES; k++) {
    double x = G.kdop[3 * k], y = G.kdop[3 * k + 1], z = G.kdop[3 * k + 2];
    double uA = -INFINITY, lA = INFINITY, uB = -INFINITY, lB = INFINITY;
    for (int i = 0; i < na; i++) {
      double lv = x * A[i][0] + y * A[i][1] + z * A[i][2];
      if (lv < lA) lA = lv;
      if (lv > uA) uA = lv;
    }
    for (int i = 0; i < nb; i++) {
      double lv = x * B[i][0] + y * B[i][1] + z * B[i][2];
      if (lv < lB) lB = lv;
      if (lv > uB) uB = lv;
    }
    if (uB < lA - d || uA < lB - d) return 0;
  }
  return 1;
}
/* [P + t0*D ; P + t1*D] (CCD.h:119-120,419-420) */
static void swept(const double (*P)[3], const double (*D)[3], double t0, double t1, double (*out)[3]) {
  for (int i = 0; i < 6; i++)
    for (int a = 0; a < 3; a++) {
      out[i][a] = P[i][a] + t0 * D[i][a];
      out[i + 6][a] = P[i][a] + t1 * D[i][a];
    }
}
static int kdop_ccd(const double (*P)[3], const double (*D)[3], const double *q, double d, double t0, double t1) {
  double A[12][3], B[1][3] = {{q[0], q[1], q[2]}};
  swept(P, D, t0, t1, A);
  return kdop_overlap(A, 12, B, 1, d);
}
/* CCD::GJKCCD (CCD.h:116-225): collide iff |v|^2 <= d^2 */
static int gjk_ccd(const double (*P)[3], const double (*D)[3], const double *q, double d, double t0, double t1) {
  double A[12][3], B[1][3] = {{q[0], q[1], q[2]}}, v[3];
  swept(P, D, t0, t1, A);
  port_gjk_witness(A, 12, B, 1, v);
  return v[0] * v[0] + v[1] * v[1] + v[2] * v[2] <= d * d;
}
static int self_kdop_ccd(const double (*P0)[3], const double (*D0)[3], const double (*P1)[3], const double (*D1)[3], double d,
                         double t0, double t1, double s0, double s1) {
  double A[12][3], B[12][3];
  swept(P0, D0, t0, t1, A); swept(P1, D1, s0, s1, B);
  return kdop_overlap(A, 12, B, 12, d);
}
/* CCD::SelfGJKCCD (CCD.h:227-352) */
static int self_gjk_ccd(const double (*P0)[3], const double (*D0)[3], const double (*P1)[3], const double (*D1)[3], double d,
                        double t0, double t1, double s0, double s1) {
  double A[12][3], B[12][3], v[3];
  swept(P0, D0, t0, t1, A); swept(P1, D1, s0, s1, B);
  port_gjk_witness(A, 12, B, 12, v);
  return v[0] * v[0] + v[1] * v[1] + v[2] * v[2] <= d * d;
}

void port_gjk(const double *A, int na, const double *B, int nb, double *v) {
  double (*a)[3] = malloc(sizeof(double[3]) * (size_t)na), (*b)[3] = malloc(sizeof(double[3]) * (size_t)nb);
  for (int i = 0; i < na; i++) for (int k = 0; k < 3; k++) a[i][k] = A[k * na + i];
  for (int i = 0; i < nb; i++) for (int k = 0; k < 3; k++) b[i][k] = B[k * nb + i];
  port_gjk_witness(a, na, b, nb, v);
  free(a); free(b);
}
int port_kdop_dcd(const double *P, const double *q, double d) {
  double A[6][3], B[1][3] = {{q[0], q[1], q[2]}};
  load6(P, A);
  return kdop_overlap(A, 6, B, 1, d);
}
int port_self_kdop_dcd(const double *P0, const double *P1, double d) {
  double A[6][3], B[6][3];
  load6(P0, A); load6(P1, B);
  return kdop_overlap(A, 6, B, 6, d);
}
int port_kdop_ccd(const double *P, const double *D, const double *q, double d, double t0, double t1) {
  double A[6][3], E[6][3];
  load6(P, A); load6(D, E);
  return kdop_ccd(A, E, q, d, t0, t1);
}
int port_gjk_ccd(const double *P, const double *D, const double *q, double d, double t0, double t1) {
  double A[6][3], E[6][3];
  load6(P, A); load6(D, E);
  return gjk_ccd(A, E, q, d, t0, t1);
}
int port_self_kdop_ccd(const double *P0, const double *D0, const double *P1, const double *D1, double d, double t0, double t1,
                       double s0, double s1) {
  double A[6][3], E[6][3], B[6][3], F[6][3];
  load6(P0, A); load6(D0, E); load6(P1, B); load6(D1, F);
  return self_kdop_ccd(A, E, B, F, d, t0, t1, s0, s1);
}
int port_self_gjk_ccd(const double *P0, const double *D0, const double *P1, const double *D1, double d, double t0, double t1,
                      double s0, double s1) {
  double A[6][3], E[6][3], B[6][3], F[6][3];
  load6(P0, A); load6(D0, E); load6(P1, B); load6(D1, F);
  return self_gjk_ccd(A, E, B, F, d, t0, t1, s0, s1);
}

/* ---- separating planes: Separate::opengjk (HighOrderCCD/Separate.h:18-163), ::selfgjk (:165-304),
 *      Optimal_plane::optimal_d (HighOrderCCD/Optimal_plane.h:13-71) ----------------------------------------------------- */
static int plane_point(const double (*P)[3], const double *q, double distance, double *c, double *d) {
  double B[1][3] = {{q[0], q[1], q[2]}};
  port_gjk_witness(P, 6, B, 1, c);
  double cn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);       /* Vector3d::norm() */
  if (cn > distance) return 0;                                     /* :111-113 */
  c[0] /= cn; c[1] /= cn; c[2] /= cn;
  double d0 = -c[0] * q[0] - c[1] * q[1] - c[2] * q[2];            /* :137-151 */
  *d = d0 - G.offset;
  return 1;
}
static int plane_hulls(const double (*P0)[3], const double (*P1)[3], double distance, double *c, double *d) {
  port_gjk_witness(P0, 6, P1, 6, c);
  double cn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (cn > distance) return 0;
  c[0] /= cn; c[1] /= cn; c[2] /= cn;
  double d0 = INFINITY, d1 = -INFINITY;                            /* :265-286 */
  for (int i = 0; i < 6; i++) {
    double t = -(c[0] * P1[i][0] + (c[1] * P1[i][1] + c[2] * P1[i][2]));   /* c.dot(row): x0 + (x1 + x2) */
    if (d0 > t) d0 = t;
  }
  for (int i = 0; i < 6; i++) {
    double t = -(c[0] * P0[i][0] + (c[1] * P0[i][1] + c[2] * P0[i][2]));
    if (d1 < t) d1 = t;
  }
  *d = 0.5 * (d0 + d1);
  return 1;
}
static void optimal_d(const double (*P0)[3], const double (*P1)[3], const double *c, double *d_io) {
  double d = *d_io;
  const double margin = G.margin, offset = G.offset;
  for (;;) {
    double g = 0, h = 0;
    for (int j = 0; j < 6; j++) {
      double dist = (P0[j][0] * c[0] + P0[j][1] * c[1] + P0[j][2] * c[2]) + d - 0.5 * offset;
      if (dist < margin) {
        double lg = log(dist / margin);
        g += -(2 * (dist - margin) * lg + (dist - margin) * (dist - margin) / dist);
        h += -(2 * lg + 4 * (dist - margin) / dist - (dist - margin) * (dist - margin) / (dist * dist));
      }
    }
    for (int j = 0; j < 6; j++) {
      double dist = -(P1[j][0] * c[0] + P1[j][1] * c[1] + P1[j][2] * c[2]) - d - 0.5 * offset;
      if (dist < margin) {
        double lg = log(dist / margin);
        g += -(-(2 * (dist - margin) * lg + (dist - margin) * (dist - margin) / dist));
        h += -(2 * lg + 4 * (dist - margin) / dist - (dist - margin) * (dist - margin) / (dist * dist));
      }
    }
    d = d + 1.0 * (-g / h);
    if (fabs(g) < 1e-2) break;
  }
  *d_io = d;
}
int port_opengjk(const double *P, const double *q, double dist, double *c, double *d) {
  double A[6][3];
  load6(P, A);
  return plane_point(A, q, dist, c, d);
}
int port_selfgjk(const double *P0, const double *P1, double dist, double *c, double *d) {
  double A[6][3], B[6][3];
  load6(P0, A); load6(P1, B);
  return plane_hulls(A, B, dist, c, d);
}
void port_optimal_d(const double *P0, const double *P1, const double *c, double *d) {
  double A[6][3], B[6][3];
  load6(P0, A); load6(P1, B);
  optimal_d(A, B, c, d);
}

/* the persistent state of Main/admmPathPlanning3D.cpp:342-351 and Main/multiPathPlanning3D.cpp:450-464, emptied */
void port_reset_persistent_planes(void) {
  free(G.is_sep); free(G.sep_cd); free(G.is_self_sep); free(G.self_sep_cd);
  size_t n = (size_t)G.n_tr * (size_t)(G.n_pts > 0 ? G.n_pts : 1), m = (size_t)G.n_tr * G.uav_num * G.uav_num;
  G.is_sep = (unsigned char *)calloc(n, 1); G.sep_cd = (double *)calloc(4 * n, sizeof(double));
  G.is_self_sep = (unsigned char *)calloc(m, 1); G.self_sep_cd = (double *)calloc(4 * m, sizeof(double));
}
long port_live_planes(unsigned *tr, unsigned *id, double *c, double *d, long cap) {
  long n = 0;
  if (!G.is_sep) return 0;
  for (int t = 0; t < G.n_tr; t++)
    for (int k = 0; k < G.n_pts; k++)
      if (G.is_sep[(size_t)t * G.n_pts + k]) {
        const double *s = G.sep_cd + 4 * ((size_t)t * G.n_pts + k);
        if (n < cap) { tr[n] = (unsigned)t; id[n] = (unsigned)k; c[3 * n] = s[0]; c[3 * n + 1] = s[1]; c[3 * n + 2] = s[2]; d[n] = s[3]; }
        n++;
      }
  return n;
}

/* ragged plane lists of one robot */
typedef struct { int n_tr; long *cnt, *cap; double **c; double **d; } Planes;
static void planes_init(Planes *p, int n_tr) {
  p->n_tr = n_tr;
  p->cnt = (long *)calloc((size_t)n_tr, sizeof(long)); p->cap = (long *)calloc((size_t)n_tr, sizeof(long));
  p->c = (double **)calloc((size_t)n_tr, sizeof(double *)); p->d = (double **)calloc((size_t)n_tr, sizeof(double *));
}
static void planes_free(Planes *p) {
  for (int t = 0; t < p->n_tr; t++) { free(p->c[t]); free(p->d[t]); }
  free(p->cnt); free(p->cap); free(p->c); free(p->d);
}
static void planes_push(Planes *p, int tr, const double *c, double d) {
  if (p->cnt[tr] == p->cap[tr]) {
    p->cap[tr] = p->cap[tr] ? 2 * p->cap[tr] : 16;
    p->c[tr] = (double *)realloc(p->c[tr], sizeof(double) * 3 * (size_t)p->cap[tr]);
    p->d[tr] = (double *)realloc(p->d[tr], sizeof(double) * (size_t)p->cap[tr]);
  }
  long k = p->cnt[tr]++;
  p->c[tr][3 * k] = c[0]; p->c[tr][3 * k + 1] = c[1]; p->c[tr][3 * k + 2] = c[2];
  p->d[tr][k] = d;
}
static void planes_from_csr(Planes *p, const unsigned *off, const double *c, const double *d) {
  planes_init(p, G.n_tr);
  for (int t = 0; t < G.n_tr; t++)
    for (unsigned k = off[t]; k < off[t + 1]; k++) planes_push(p, t, c + 3 * (size_t)k, d[k]);
}
static long planes_to_csr(const Planes *p, unsigned *off, double *c, double *d, long cap, long n0) {
  long n = n0;
  for (int t = 0; t < p->n_tr; t++) {
    for (long k = 0; k < p->cnt[t]; k++) {
      if (n < cap) { c[3 * n] = p->c[t][3 * k]; c[3 * n + 1] = p->c[t][3 * k + 1]; c[3 * n + 2] = p->c[t][3 * k + 2]; d[n] = p->d[t][k]; }
      n++;
    }
    off[t + 1] = (unsigned)n;
  }
  return n;
}

/* Optimization3D_admm::separate_plane (HighOrderCCD/Optimization/Optimization3D_admm.h:69-197); `persistent`: the
 * is_optimal_plane branches :126-145 (a pair that separated once keeps its plane) and :164-193 (every live plane of the
 * sub-segment is refined by optimal_cd and emitted in point-id order) */
static void separate_plane_mode(const double *spline, Planes *pl, int persistent) {
  const double dist = G.offset + G.margin;
  const int np = G.n_pts;
  unsigned *ids = NULL;
  long cap = 0;
  for (int tr = 0; tr < G.n_tr; tr++) {
    double P[6][3], lo[3], hi[3];
    seg_points(spline, tr, P);
    box_of(P, 6, lo, hi, 1);
    long n = box_query(lo, hi, dist, ids, cap);
    if (n > cap) { cap = n + 1024; ids = (unsigned *)realloc(ids, sizeof(unsigned) * (size_t)cap); n = box_query(lo, hi, dist, ids, cap); }
    for (long i = 0; i < n; i++) {
      double q[3] = {G.V[ids[i]], G.V[(size_t)np + ids[i]], G.V[(size_t)2 * np + ids[i]]};
      double B[1][3] = {{q[0], q[1], q[2]}}, c[3], d;
      if (!kdop_overlap(P, 6, B, 1, dist)) continue;
      if (persistent) {
        size_t at = (size_t)tr * np + ids[i];
        if (!G.is_sep[at] && plane_point(P, q, dist, c, &d)) {
          G.is_sep[at] = 1;
          G.sep_cd[4 * at] = c[0]; G.sep_cd[4 * at + 1] = c[1]; G.sep_cd[4 * at + 2] = c[2]; G.sep_cd[4 * at + 3] = d;
        }
      } else if (plane_point(P, q, dist, c, &d)) planes_push(pl, tr, c, d);
    }
    if (persistent)
      for (int ob = 0; ob < np; ob++) {
        size_t at = (size_t)tr * np + ob;
        if (!G.is_sep[at]) continue;
        double q[3] = {G.V[ob], G.V[(size_t)np + ob], G.V[(size_t)2 * np + ob]};
        double *s = G.sep_cd + 4 * at;
        port_optimal_cd_impl(P, q, s, s + 3);
        planes_push(pl, tr, s, s[3]);
      }
  }
  free(ids);
}
static void separate_plane(const double *spline, Planes *pl) {
  if (G.optimal_plane && !G.is_sep) port_reset_persistent_planes();
  separate_plane_mode(spline, pl, G.optimal_plane);
}
/* Optimization3D_multi::separate_plane (Optimization3D_multi.h:176-235) has no is_optimal_plane branch */
static void separate_plane_multi(const double *spline, Planes *pl) { separate_plane_mode(spline, pl, 0); }
long port_separate_plane(const double *spline, unsigned *off, double *c, double *d, long cap) {
  Planes pl;
  planes_init(&pl, G.n_tr);
  separate_plane(spline, &pl);
  off[0] = 0;
  long n = planes_to_csr(&pl, off, c, d, cap, 0);
  planes_free(&pl);
  return n;
}

/* Optimization3D_multi::separate_self (Optimization/Optimization3D_multi.h:237-342): appends to the lists of both robots of
 * every accepted pair; is_optimal_plane: :271-284 (first separation is kept, no optimal_d) and :309-339 (all live pairs of
 * the slot refined by self_optimal_cd, emitted in (p0, p1) order) */
static void separate_self(const double *splines, int u, Planes *pls) {
  const double dist = G.offset + 2 * G.margin;
  if (G.optimal_plane && !G.is_self_sep) port_reset_persistent_planes();
  const size_t ns = (size_t)3 * G.T;
  double (*Pl)[6][3] = malloc(sizeof(double[6][3]) * (size_t)u);
  double *lo = (double *)malloc(sizeof(double) * 3 * (size_t)u), *hi = (double *)malloc(sizeof(double) * 3 * (size_t)u);
  for (int tr = 0; tr < G.n_tr; tr++) {
    for (int i = 0; i < u; i++) { seg_points(splines + ns * i, tr, Pl[i]); box_of(Pl[i], 6, lo + 3 * i, hi + 3 * i, 1); }
    for (int p0 = 0; p0 < u; p0++)
      for (int p1 = p0 + 1; p1 < u; p1++) {
        if (!boxes_within(lo + 3 * p0, hi + 3 * p0, lo + 3 * p1, hi + 3 * p1, dist)) continue;
        if (!kdop_overlap(Pl[p0], 6, Pl[p1], 6, dist)) continue;
        double c[3], d;
        if (G.optimal_plane) {
          size_t at = ((size_t)tr * u + p0) * u + p1;
          if (!G.is_self_sep[at] && plane_hulls(Pl[p0], Pl[p1], dist, c, &d)) {
            G.is_self_sep[at] = 1;
            G.self_sep_cd[4 * at] = c[0]; G.self_sep_cd[4 * at + 1] = c[1]; G.self_sep_cd[4 * at + 2] = c[2]; G.self_sep_cd[4 * at + 3] = d;
          }
          continue;
        }
        if (!plane_hulls(Pl[p0], Pl[p1], dist, c, &d)) continue;
        optimal_d(Pl[p0], Pl[p1], c, &d);
        double cm[3] = {-c[0], -c[1], -c[2]};
        planes_push(&pls[p0], tr, c, d - 0.5 * G.offset);
        planes_push(&pls[p1], tr, cm, -d - 0.5 * G.offset);
      }
    if (G.optimal_plane)
      for (int p0 = 0; p0 < u; p0++)
        for (int p1 = p0 + 1; p1 < u; p1++) {
          size_t at = ((size_t)tr * u + p0) * u + p1;
          if (!G.is_self_sep[at]) continue;
          double *s = G.self_sep_cd + 4 * at;
          port_self_optimal_cd_impl(Pl[p0], Pl[p1], s, s + 3);
          double cm[3] = {-s[0], -s[1], -s[2]};
          planes_push(&pls[p0], tr, s, s[3] - 0.5 * G.offset);
          planes_push(&pls[p1], tr, cm, -s[3] - 0.5 * G.offset);
        }
  }
  free(Pl); free(lo); free(hi);
}
long port_separate_self(const double *splines, int u, unsigned *off, double *c, double *d, long cap) {
  Planes *pls = (Planes *)malloc(sizeof(Planes) * (size_t)u);
  for (int i = 0; i < u; i++) planes_init(&pls[i], G.n_tr);
  separate_self(splines, u, pls);
  long n = 0;
  off[0] = 0;
  for (int i = 0; i < u; i++) { n = planes_to_csr(&pls[i], off + (size_t)i * G.n_tr, c, d, cap, n); planes_free(&pls[i]); }
  free(pls);
  return n;
}

/* ---- energies: HighOrderCCD/Energy_admm.h ---------------------------------------------------------------------------------- */
/* plane_barrier_energy :46-96 */
static double plane_barrier_energy(const double *spline, const Planes *pl) {
  double energy = 0;
  const double margin = G.margin;
  for (int tr = 0; tr < G.n_tr; tr++) {
    if (pl->cnt[tr] == 0) continue;
    const double w = G.weight[tr];
    double P[6][3];
    seg_points(spline, tr, P);
    for (long k = 0; k < pl->cnt[tr]; k++) {
      const double *c = pl->c[tr] + 3 * k;
      for (int j = 0; j <= ORDER; j++) {
        double d = P[j][0] * c[0] + P[j][1] * c[1] + P[j][2] * c[2] + pl->d[tr][k];
        if (d <= 0) return INFINITY;
        if (d < margin) energy += -w * (d - margin) * (d - margin) * log(d / margin);
      }
    }
  }
  return energy;
}
/* bound_energy :98-170 */
static double bound_energy(const double *spline, double piece_time) {
  double energy = 0;
  const double margin = G.margin;
  for (int tr = 0; tr < G.n_tr; tr++) {
    const double w = G.weight[tr];
    double P[6][3];
    seg_points(spline, tr, P);
    for (int j = 0; j < ORDER; j++) {
      double v[3];
      for (int a = 0; a < 3; a++) v[a] = ORDER * (P[j + 1][a] - P[j][a]);
      double d = G.vel_limit - sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / (w * piece_time);
      if (d <= 0) return INFINITY;
      if (d < margin) energy += -w * (d - margin) * (d - margin) * log(d / margin);
    }
    for (int j = 0; j < ORDER - 1; j++) {
      double a3[3];
      for (int a = 0; a < 3; a++) a3[a] = (ORDER * (ORDER - 1)) * (P[j + 2][a] - 2 * P[j + 1][a] + P[j][a]);
      double d = G.acc_limit - sqrt(a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2]) / (w * w * piece_time * piece_time);
      if (d <= 0) return INFINITY;
      if (d < margin) energy += -w * (d - margin) * (d - margin) * log(d / margin);
    }
  }
  return energy;
}
/* c_spline = convert_list[sp] * spline.block<6,3>(3 sp, 0): out[r + 6*axis] */
static void convert_piece(const double *spline, int sp, double *out) {
  const double *Cv = G.convert + 36 * sp;
  for (int ax = 0; ax < 3; ax++)
    for (int r = 0; r < 6; r++) {
      double acc = 0;
      for (int k = 0; k < 6; k++) acc += M6(Cv, r, k) * spline[(size_t)ax * G.T + 3 * sp + k];
      out[r + 6 * ax] = acc;
    }
}
/* spline_energy :16-44 */
static double spline_energy(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                            const double *p_lambda, const double *t_lambda, const Planes *pl) {
  const int Pn = G.piece_num;
  double energy = G.lambda * plane_barrier_energy(spline, pl) + G.lambda * bound_energy(spline, piece_time);
  for (int sp = 0; sp < Pn; sp++) {
    double cs[18], pd[18], sq = 0;
    convert_piece(spline, sp, cs);
    for (int ax = 0; ax < 3; ax++)
      for (int r = 0; r < 6; r++) { pd[r + 6 * ax] = cs[r + 6 * ax] - p_slack[(size_t)ax * 6 * Pn + 6 * sp + r]; }
    for (int i = 0; i < 18; i++) sq += pd[i] * pd[i];
    energy += G.mu / 2.0 * sq;
    energy += G.mu / 2.0 * pow(piece_time - t_slack[sp], 2);
    for (int ax = 0; ax < 3; ax++) {
      double dot = 0;
      for (int r = 0; r < 6; r++) dot += p_lambda[(size_t)ax * 6 * Pn + 6 * sp + r] * pd[r + 6 * ax];
      energy += dot;
    }
    energy += t_lambda[sp] * (piece_time - t_slack[sp]);
  }
  return energy;
}
/* dynamic_energy :199-215, slack_energy :172-190.  6x3 col-major blocks */
static double dynamic_energy(const double *p_part, double t_part) {
  double energy = 0;
  for (int ax = 0; ax < 3; ax++) {
    double quad = 0;
    for (int r = 0; r < 6; r++) {
      double mx = 0;
      for (int k = 0; k < 6; k++) mx += M6(G.mdyn, r, k) * p_part[k + 6 * ax];
      quad += p_part[r + 6 * ax] * mx;
    }
    energy += G.ks / pow(t_part, 5) * 0.5 * quad;
  }
  energy += G.kt * pow(t_part, 1.1);
  return energy;
}
static double slack_energy(const double *c_spline, double piece_time, const double *p_part, double t_part, const double *p_lambda,
                           double t_lambda) {
  double energy = dynamic_energy(p_part, t_part), sq = 0;
  for (int i = 0; i < 18; i++) sq += (c_spline[i] - p_part[i]) * (c_spline[i] - p_part[i]);
  energy += G.mu / 2.0 * sq;
  energy += G.mu / 2.0 * (piece_time - t_part) * (piece_time - t_part);
  for (int ax = 0; ax < 3; ax++) {
    double dot = 0;
    for (int r = 0; r < 6; r++) dot += p_lambda[r + 6 * ax] * (c_spline[r + 6 * ax] - p_part[r + 6 * ax]);
    energy += dot;
  }
  energy += t_lambda * (piece_time - t_part);
  return energy;
}

double port_plane_barrier_energy(const double *spline, const unsigned *off, const double *c, const double *d) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  double e = plane_barrier_energy(spline, &pl);
  planes_free(&pl);
  return e;
}
double port_bound_energy(const double *spline, double piece_time) { return bound_energy(spline, piece_time); }
double port_spline_energy(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                          const double *p_lambda, const double *t_lambda, const unsigned *off, const double *c, const double *d) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  double e = spline_energy(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl);
  planes_free(&pl);
  return e;
}
double port_slack_energy(const double *c_spline, double piece_time, const double *p_part, double t_part, const double *p_lambda,
                         double t_lambda) {
  return slack_energy(c_spline, piece_time, p_part, t_part, p_lambda, t_lambda);
}

/* ---- gradients: HighOrderCCD/Gradient_admm.h.  Piece coordinates: index 3*m + k (control point m, axis k) ------------------ */
/* local_plane_barrier_gradient :331-407: d_x = A_list[tr][j] * c = kron(basis.row(j), I3)^T c */
static void local_plane_barrier_gradient(int tr, const double *spline, const Planes *pl, double *grad, double *hess) {
  const double *B = G.basis + (size_t)36 * tr;
  const double w = G.weight[tr], margin = G.margin;
  double P[6][3];
  seg_points(spline, tr, P);
  memset(grad, 0, sizeof(double) * 18);
  memset(hess, 0, sizeof(double) * 324);
  for (int j = 0; j <= ORDER; j++)
    for (long k = 0; k < pl->cnt[tr]; k++) {
      const double *c = pl->c[tr] + 3 * k;
      double d = P[j][0] * c[0] + P[j][1] * c[1] + P[j][2] * c[2] + pl->d[tr][k];
      if (d < margin) {
        double dx[18];
        for (int m = 0; m < 6; m++) for (int a = 0; a < 3; a++) dx[3 * m + a] = M6(B, j, m) * c[a];
        double e1 = -w * (2 * (d - margin) * log(d / margin) + (d - margin) * (d - margin) / d);
        double e2 = -w * (2 * log(d / margin) + 4 * (d - margin) / d - (d - margin) * (d - margin) / (d * d));
        for (int i = 0; i < 18; i++) grad[i] += e1 * dx[i];
        for (int c2 = 0; c2 < 18; c2++) for (int r = 0; r < 18; r++) hess[r + 18 * c2] += e2 * dx[r] * dx[c2];
      }
    }
}
/* local_bound_gradient :409-572 */
static void local_bound_gradient(int tr, const double *spline, double t, double *grad, double *hess, double *g_t, double *h_t,
                                 double *partgrad) {
  const double *B = G.basis + (size_t)36 * tr;
  const double w = G.weight[tr], margin = G.margin;
  double P[6][3];
  seg_points(spline, tr, P);
  memset(grad, 0, sizeof(double) * 18);
  memset(hess, 0, sizeof(double) * 324);
  memset(partgrad, 0, sizeof(double) * 18);
  *g_t = 0; *h_t = 0;
  for (int pass = 0; pass < 2; pass++) {                 /* 0: velocity (:445-507), 1: acceleration (:509-570) */
    const int nj = pass == 0 ? ORDER : ORDER - 1;
    for (int j = 0; j < nj; j++) {
      double p_[3], a[6];                                /* a = row of A_vel / A_acc restricted to one axis */
      for (int k = 0; k < 3; k++) p_[k] = pass == 0 ? P[j + 1][k] - P[j][k] : P[j + 2][k] - 2 * P[j + 1][k] + P[j][k];
      for (int m = 0; m < 6; m++) a[m] = pass == 0 ? M6(B, j + 1, m) - M6(B, j, m) : M6(B, j + 2, m) - 2 * M6(B, j + 1, m) + M6(B, j, m);
      double dn = sqrt(p_[0] * p_[0] + p_[1] * p_[1] + p_[2] * p_[2]);
      double val = pass == 0 ? ORDER * dn / w : ORDER * (ORDER - 1) * dn / (w * w);
      double d = pass == 0 ? G.vel_limit - val / t : G.acc_limit - val / (t * t);
      if (!(d < margin)) continue;
      double e1 = -w * (2 * (d - margin) * log(d / margin) + (d - margin) * (d - margin) / d);
      double e2 = -w * (2 * log(d / margin) + 4 * (d - margin) / d - (d - margin) * (d - margin) / (d * d));
      double coef, e3;
      if (pass == 0) {
        *g_t += e1 * val / pow(t, 2);
        *h_t += -2 * e1 * val / pow(t, 3) + e2 * val * val / pow(t, 4);
        coef = -ORDER / (w * t);
        e3 = -e1 / t + e2 * (G.vel_limit - d) / t;
      } else {
        *g_t += 2 * e1 * val / pow(t, 3);
        *h_t += -6 * e1 * val / pow(t, 4) + 4 * e2 * val * val / pow(t, 6);
        coef = -ORDER * (ORDER - 1) / pow(w * t, 2);
        e3 = -2 * e1 / t + 2 * e2 * (G.acc_limit - d) / t;
      }
      double dp[3], hp[3][3];
      for (int k = 0; k < 3; k++) dp[k] = coef * p_[k] / dn;
      for (int r = 0; r < 3; r++)
        for (int s = 0; s < 3; s++) hp[r][s] = coef * ((r == s ? 1.0 / dn : 0.0) - p_[r] * p_[s] / pow(dn, 3));
      double dx[18];
      for (int m = 0; m < 6; m++) for (int k = 0; k < 3; k++) dx[3 * m + k] = dp[k] * a[m];
      for (int i = 0; i < 18; i++) { grad[i] += e1 * dx[i]; partgrad[i] += e3 * dx[i]; }
      for (int m2 = 0; m2 < 6; m2++) for (int k2 = 0; k2 < 3; k2++)
        for (int m1 = 0; m1 < 6; m1++) for (int k1 = 0; k1 < 3; k1++)
          hess[(3 * m1 + k1) + 18 * (3 * m2 + k2)] += e2 * dx[3 * m1 + k1] * dx[3 * m2 + k2] + e1 * a[m1] * hp[k1][k2] * a[m2];
    }
  }
}
/* local_spline_gradient :67-164: g0[19], h0[19x19 col-major] */
static void local_spline_gradient(const double *spline, double t, const double *p_slack, const double *t_slack,
                                  const double *p_lambda, const double *t_lambda, const Planes *pl, int sp, double *g0, double *h0) {
  const int Pn = G.piece_num, num = 18, ld = 19;
  memset(g0, 0, sizeof(double) * 19);
  memset(h0, 0, sizeof(double) * 361);
  double g[18], h[324], pg[18], g_t, h_t;
  for (int i = 0; i < G.res; i++) {
    int tr = sp * G.res + i;
    if (pl->cnt[tr] == 0) continue;
    local_plane_barrier_gradient(tr, spline, pl, g, h);
    for (int r = 0; r < num; r++) g0[r] += g[r];
    for (int c = 0; c < num; c++) for (int r = 0; r < num; r++) h0[r + ld * c] += h[r + 18 * c];
  }
  for (int i = 0; i < G.res; i++) {
    int tr = sp * G.res + i;
    local_bound_gradient(tr, spline, t, g, h, &g_t, &h_t, pg);
    for (int r = 0; r < num; r++) g0[r] += g[r];
    g0[num] += g_t;
    for (int c = 0; c < num; c++) for (int r = 0; r < num; r++) h0[r + ld * c] += h[r + 18 * c];
    for (int r = 0; r < num; r++) { h0[r + ld * num] += pg[r]; h0[num + ld * r] += pg[r]; }
    h0[num + ld * num] += h_t;
  }
  for (int i = 0; i < 19; i++) g0[i] *= G.lambda;
  for (int i = 0; i < 361; i++) h0[i] *= G.lambda;
  /* consensus terms :132-163 */
  const double *Cv = G.convert + 36 * sp;
  double cs[18];
  convert_piece(spline, sp, cs);
  for (int m = 0; m < 6; m++)
    for (int k = 0; k < 3; k++) {
      double x1 = 0, x2 = 0;
      for (int r = 0; r < 6; r++) {
        size_t s = (size_t)k * 6 * Pn + 6 * sp + r;
        x1 += M6(Cv, r, m) * (cs[r + 6 * k] - p_slack[s]);
        x2 += M6(Cv, r, m) * p_lambda[s];
      }
      g0[3 * m + k] += G.mu * x1 + x2;
    }
  for (int m1 = 0; m1 < 6; m1++)
    for (int m2 = 0; m2 < 6; m2++) {
      double ctc = 0;
      for (int r = 0; r < 6; r++) ctc += M6(Cv, r, m1) * M6(Cv, r, m2);
      for (int k = 0; k < 3; k++) h0[(3 * m1 + k) + ld * (3 * m2 + k)] += G.mu * ctc;
    }
  g0[num] += G.mu * (t - t_slack[sp]) + t_lambda[sp];
  h0[num + ld * num] += G.mu;
}
/* global_spline_gradient :13-65: per-piece PSD projection, scatter-add into the dense (3T+1) system */
static void global_spline_gradient(const double *spline, double t, const double *p_slack, const double *t_slack,
                                   const double *p_lambda, const double *t_lambda, const Planes *pl, double *grad, double *hess) {
  const int n = 3 * G.T, ld = n + 1, num = 18;
  memset(grad, 0, sizeof(double) * (size_t)ld);
  memset(hess, 0, sizeof(double) * (size_t)ld * ld);
  for (int sp = 0; sp < G.piece_num; sp++) {
    double g0[19], h0[361], L[361];
    local_spline_gradient(spline, t, p_slack, t_slack, p_lambda, t_lambda, pl, sp, g0, h0);
    if (!port_llt(h0, L, 19)) {
      double ev = port_min_eig(h0, 19);
      if (ev < 0) for (int i = 0; i < 19; i++) h0[i + 19 * i] = h0[i + 19 * i] - ev * 1.0 + 0.01 * 1.0;
    }
    const int base = 3 * (sp * (ORDER - 2));
    for (int i = 0; i < num; i++) grad[base + i] += g0[i];
    for (int c = 0; c < num; c++) for (int r = 0; r < num; r++) hess[(base + r) + (size_t)ld * (base + c)] += h0[r + 19 * c];
    grad[n] += g0[num];
    hess[n + (size_t)ld * n] += h0[num + 19 * num];
    for (int r = 0; r < num; r++) {
      hess[(base + r) + (size_t)ld * n] += h0[r + 19 * num];
      hess[n + (size_t)ld * (base + r)] += h0[num + 19 * r];
    }
  }
}

void port_local_spline_gradient(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                                const double *p_lambda, const double *t_lambda, const unsigned *off, const double *c,
                                const double *d, int sp_id, double *g0, double *h0) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  local_spline_gradient(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl, sp_id, g0, h0);
  planes_free(&pl);
}
void port_global_spline_gradient(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                                 const double *p_lambda, const double *t_lambda, const unsigned *off, const double *c,
                                 const double *d, double *grad, double *hess) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  global_spline_gradient(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl, grad, hess);
  planes_free(&pl);
}

/* ---- Newton direction: Optimization3D_admm::spline_descent_direction (Optimization3D_admm.h:400-503; SimplicialLLT) and
 *      Optimization3D_multi::spline_descent_direction (Optimization3D_multi.h:659-752; dense LLT + eigenvalue shift) --------- */
static void descent_direction(const double *spline, double t, const double *p_slack, const double *t_slack, const double *p_lambda,
                              const double *t_lambda, const Planes *pl, int multi, double *direction, double *t_direction,
                              double *wolfe, double *gnorm) {
  const int T = G.T, n = 3 * T, ld = n + 1, m = 3 * (T - 4), mm = m + 1;
  double *grad = (double *)malloc(sizeof(double) * (size_t)ld), *hess = (double *)malloc(sizeof(double) * (size_t)ld * ld);
  global_spline_gradient(spline, t, p_slack, t_slack, p_lambda, t_lambda, pl, grad, hess);
  double *g0 = (double *)malloc(sizeof(double) * (size_t)mm), *h0 = (double *)malloc(sizeof(double) * (size_t)mm * mm);
  double *L = (double *)malloc(sizeof(double) * (size_t)mm * mm), *x = (double *)malloc(sizeof(double) * (size_t)mm);
  for (int i = 0; i < m; i++) g0[i] = grad[6 + i];
  g0[m] = grad[n];
  for (int c = 0; c < m; c++) for (int r = 0; r < m; r++) h0[r + (size_t)mm * c] = hess[(6 + r) + (size_t)ld * (6 + c)];
  for (int r = 0; r < m; r++) { h0[r + (size_t)mm * m] = hess[(6 + r) + (size_t)ld * n]; h0[m + (size_t)mm * r] = hess[(6 + r) + (size_t)ld * n]; }
  h0[m + (size_t)mm * m] = hess[n + (size_t)ld * n];
  int ok = port_llt(h0, L, mm);
  if (!ok && multi) {
    double ev = port_min_eig(h0, mm);
    if (ev < 0) for (int i = 0; i < mm; i++) h0[i + (size_t)mm * i] = h0[i + (size_t)mm * i] - ev * 1.0 + 0.01 * 1.0;
    port_llt(h0, L, mm);
  }
  for (int i = 0; i < mm; i++) x[i] = g0[i];
  port_llt_solve(L, mm, x);
  double wl = 0, gn = 0;
  for (int i = 0; i < mm; i++) { x[i] = -x[i]; wl += x[i] * g0[i]; gn += g0[i] * g0[i]; }
  *wolfe = -wl;
  *gnorm = sqrt(gn);
  memset(direction, 0, sizeof(double) * (size_t)n);
  for (int p = 0; p < T - 4; p++) for (int k = 0; k < 3; k++) direction[(size_t)k * T + 2 + p] = x[3 * p + k];
  *t_direction = x[m];
  free(grad); free(hess); free(g0); free(h0); free(L); free(x);
}
void port_descent_direction(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                            const double *p_lambda, const double *t_lambda, const unsigned *off, const double *c, const double *d,
                            double *direction, double *t_direction, double *wolfe_out, double *gnorm_out) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  descent_direction(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl, 0, direction, t_direction, wolfe_out, gnorm_out);
  planes_free(&pl);
}
void port_descent_direction_multi(const double *spline, double piece_time, const double *p_slack, const double *t_slack,
                                  const double *p_lambda, const double *t_lambda, const unsigned *off, const double *c,
                                  const double *d, double *direction, double *t_direction, double *wolfe_out, double *gnorm_add) {
  Planes pl;
  planes_from_csr(&pl, off, c, d);
  descent_direction(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl, 1, direction, t_direction, wolfe_out, gnorm_add);
  planes_free(&pl);
}

/* ---- CCD step bounds: HighOrderCCD/Step.h ---------------------------------------------------------------------------------- */
/* Step::position_step :21-110: `step` is carried across pairs */
static double position_step(const double *spline, const double *direction) {
  double step = 1.0;
  const int np = G.n_pts;
  unsigned *ids = NULL;
  long cap = 0;
  for (int tr = 0; tr < G.n_tr; tr++) {
    const int sp = tr / G.res, T = G.T;
    const double *B = G.basis + (size_t)36 * tr;
    double P[6][3], D[6][3], Q[6][3], lo[3], hi[3];
    seg_points(spline, tr, P);
    seg_points_moved(spline, direction, tr, Q);
    for (int ax = 0; ax < 3; ax++)
      for (int j = 0; j < 6; j++) {
        double acc = 0;
        for (int k = 0; k < 6; k++) acc += M6(B, j, k) * direction[(size_t)ax * T + 3 * sp + k];
        D[j][ax] = acc;
      }
    box_of(P, 6, lo, hi, 1);
    box_of(Q, 6, lo, hi, 0);
    long n = box_query(lo, hi, G.offset, ids, cap);
    if (n > cap) { cap = n + 1024; ids = (unsigned *)realloc(ids, sizeof(unsigned) * (size_t)cap); n = box_query(lo, hi, G.offset, ids, cap); }
    for (long i = 0; i < n; i++) {
      double q[3] = {G.V[ids[i]], G.V[(size_t)np + ids[i]], G.V[(size_t)2 * np + ids[i]]};
      int hit = kdop_ccd(P, D, q, G.offset, 0, step);
      while (hit) {
        hit = gjk_ccd(P, D, q, G.offset, 0, step);
        if (hit) step *= 0.8;
      }
    }
  }
  free(ids);
  return step;
}
double port_position_step(const double *spline, const double *direction) { return position_step(spline, direction); }

/* Step::self_step :184-256 (coupled = 0: one step per robot) and Step::couple_self_step :112-182 (coupled = 1: one shared step) */
static void self_steps(const double *splines, const double *directions, int u, int coupled, double *steps) {
  const size_t ns = (size_t)3 * G.T;
  double (*Pl)[6][3] = malloc(sizeof(double[6][3]) * (size_t)u), (*Dl)[6][3] = malloc(sizeof(double[6][3]) * (size_t)u);
  double *lo = (double *)malloc(sizeof(double) * 3 * (size_t)u), *hi = (double *)malloc(sizeof(double) * 3 * (size_t)u);
  double *zero = (double *)calloc(ns, sizeof(double));
  for (int i = 0; i < u; i++) steps[i] = 1.0;
  double shared = 1.0;
  for (int tr = 0; tr < G.n_tr; tr++) {
    for (int i = 0; i < u; i++) {
      double Q[6][3];
      seg_points(splines + ns * i, tr, Pl[i]);
      seg_points(directions + ns * i, tr, Dl[i]);          /* D = basis * bz_d */
      for (int j = 0; j < 6; j++) for (int a = 0; a < 3; a++) Q[j][a] = Pl[i][j][a] + Dl[i][j][a];   /* BVH.cpp:300-321 */
      box_of(Pl[i], 6, lo + 3 * i, hi + 3 * i, 1);
      box_of(Q, 6, lo + 3 * i, hi + 3 * i, 0);
    }
    for (int p0 = 0; p0 < u; p0++)
      for (int p1 = p0 + 1; p1 < u; p1++) {
        if (!boxes_within(lo + 3 * p0, hi + 3 * p0, lo + 3 * p1, hi + 3 * p1, G.offset)) continue;
        double s0 = coupled ? shared : steps[p0], s1 = coupled ? shared : steps[p1];
        int hit = self_kdop_ccd(Pl[p0], Dl[p0], Pl[p1], Dl[p1], G.offset, 0, s0, 0, s1);
        while (hit) {
          hit = self_gjk_ccd(Pl[p0], Dl[p0], Pl[p1], Dl[p1], G.offset, 0, s0, 0, s1);
          if (hit) { s0 *= 0.8; s1 *= 0.8; }
        }
        if (coupled) shared = s0;
        else { steps[p0] = s0; steps[p1] = s1; }
      }
  }
  if (coupled) steps[0] = shared;
  free(Pl); free(Dl); free(lo); free(hi); free(zero);
}
void port_self_step(const double *splines, const double *directions, int u, double *steps) { self_steps(splines, directions, u, 0, steps); }
double port_couple_self_step(const double *splines, const double *directions, int u) {
  double *st = (double *)malloc(sizeof(double) * (size_t)(u > 0 ? u : 1));
  self_steps(splines, directions, u, 1, st);
  double s = st[0];
  free(st);
  return s;
}

/* ---- slack / dual update: Optimization3D_admm::update_slack_lambda (Optimization3D_admm.h:231-398) with
 *      Gradient_admm::slack_gradient (Gradient_admm.h:574-622) and ::dynamic_gradient (:633-671) --------------------------------- */
static void slack_gradient(const double *c_spline, double piece_time, const double *p_part, double t_part, const double *p_lambda,
                           double t_lambda, double *grad, double *hess) {
  const int n = 18, ld = 19;
  const double c5 = G.ks / pow(t_part, 5);
  memset(grad, 0, sizeof(double) * 19);
  memset(hess, 0, sizeof(double) * 361);
  double dyn = 0;
  for (int k = 0; k < 3; k++) {
    double quad = 0;
    for (int m = 0; m < 6; m++) {
      double mx = 0;
      for (int s = 0; s < 6; s++) mx += M6(G.mdyn, m, s) * p_part[s + 6 * k];
      grad[3 * m + k] = c5 * mx;
      quad += p_part[m + 6 * k] * mx;
    }
    dyn += c5 * 0.5 * quad;
  }
  for (int m1 = 0; m1 < 6; m1++) for (int m2 = 0; m2 < 6; m2++) for (int k = 0; k < 3; k++) hess[(3 * m1 + k) + ld * (3 * m2 + k)] = c5 * M6(G.mdyn, m1, m2);
  double g_t = -5 * dyn / t_part + G.kt * 1.1 * pow(t_part, 0.1);
  double h_t = 5 * 6 * dyn / (t_part * t_part) + G.kt * 0.11 * pow(t_part, -0.9);
  for (int i = 0; i < n; i++) { double pg = -5 * grad[i] / t_part; hess[i + ld * n] = pg; hess[n + ld * i] = pg; }
  for (int m = 0; m < 6; m++) for (int k = 0; k < 3; k++) grad[3 * m + k] += G.mu * (p_part[m + 6 * k] - c_spline[m + 6 * k]) - p_lambda[m + 6 * k];
  for (int i = 0; i < n; i++) hess[i + ld * i] += G.mu;
  g_t += G.mu * (t_part - piece_time) - t_lambda;
  h_t += G.mu;
  grad[n] = g_t;
  hess[n + ld * n] = h_t;
}
static void update_slack_lambda(const double *spline, double piece_time, double *p_slack, double *t_slack, double *p_lambda,
                                double *t_lambda) {
  const int Pn = G.piece_num;
  for (int sp = 0; sp < Pn; sp++) {
    double cs[18], pp[18], pl[18], grad[19], hess[361];
    convert_piece(spline, sp, cs);
    for (int ax = 0; ax < 3; ax++) for (int r = 0; r < 6; r++) {
      pp[r + 6 * ax] = p_slack[(size_t)ax * 6 * Pn + 6 * sp + r];
      pl[r + 6 * ax] = p_lambda[(size_t)ax * 6 * Pn + 6 * sp + r];
    }
    double t_part = t_slack[sp], tl = t_lambda[sp];
    slack_gradient(cs, piece_time, pp, t_part, pl, tl, grad, hess);
    int off = 0, tn = 6;                                   /* first piece drops points 0,1; last piece drops 4,5 */
    if (sp == 0) { off = 6; tn = 4; }
    else if (sp == Pn - 1) { off = 0; tn = 4; }
    const int n = 3 * tn + 1;
    double g0[19], h0[361], L[361], x[19];
#define GI(i) ((i) < 3 * tn ? off + (i) : 18)
    for (int i = 0; i < n; i++) g0[i] = grad[GI(i)];
    for (int c = 0; c < n; c++) for (int r = 0; r < n; r++) h0[r + n * c] = hess[GI(r) + 19 * GI(c)];
#undef GI
    if (!port_llt(h0, L, n)) {
      double ev = port_min_eig(h0, n);
      if (ev < 0) for (int i = 0; i < n; i++) h0[i + n * i] = h0[i + n * i] - ev * 1.0 + 0.01 * 1.0;
      port_llt(h0, L, n);
    }
    for (int i = 0; i < n; i++) x[i] = g0[i];
    port_llt_solve(L, n, x);
    double wl = 0;
    for (int i = 0; i < n; i++) { x[i] = -x[i]; wl += x[i] * g0[i]; }
    G.wolfe = -wl;
    double dir[18];
    memset(dir, 0, sizeof(dir));
    for (int p = 0; p < tn; p++) for (int k = 0; k < 3; k++) dir[(sp == 0 ? 2 : 0) + p + 6 * k] = x[3 * p + k];
    const double tdir = x[3 * tn];
    double step = 1.0;
    if (t_part + step * tdir <= 0) step = -0.95 * t_part / tdir;
    const double e = slack_energy(cs, piece_time, pp, t_part, pl, tl);
    const double init_time = t_part;
    t_part = init_time + step * tdir;
    double pn[18];
    for (;;) {
      for (int i = 0; i < 18; i++) pn[i] = pp[i] + step * dir[i];
      if (!(e - 1e-4 * G.wolfe * step < slack_energy(cs, piece_time, pn, t_part, pl, tl))) break;
      step *= 0.8;
      t_part = init_time + step * tdir;
    }
    for (int ax = 0; ax < 3; ax++) for (int r = 0; r < 6; r++) {
      size_t s = (size_t)ax * 6 * Pn + 6 * sp + r;
      double v = pp[r + 6 * ax] + step * dir[r + 6 * ax];
      p_slack[s] = v;
      p_lambda[s] += G.mu * (cs[r + 6 * ax] - v);
    }
    t_slack[sp] = t_part;
    t_lambda[sp] += G.mu * (piece_time - t_part);
  }
}
void port_update_slack_lambda(const double *spline, double piece_time, double *p_slack, double *t_slack, double *p_lambda,
                              double *t_lambda) {
  update_slack_lambda(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
}

/* ---- line search: Optimization3D_admm::spline_line_search (Optimization3D_admm.h:505-557) / Optimization3D_multi :754-811 --- */
static void line_search(double *spline, const double *direction, double *piece_time, double t_direction, const double *p_slack,
                        const double *t_slack, const double *p_lambda, const double *t_lambda, const Planes *pl, double step,
                        double wolfe) {
  const size_t n = (size_t)3 * G.T;
  if (*piece_time + step * t_direction <= 0) step = -0.95 * *piece_time / t_direction;
  const double e = spline_energy(spline, *piece_time, p_slack, t_slack, p_lambda, t_lambda, pl);
  const double init_time = *piece_time;
  double *trial = (double *)malloc(sizeof(double) * n);
  double t = init_time + step * t_direction;
  for (int guard = 0; guard < 2000; guard++) {
    for (size_t i = 0; i < n; i++) trial[i] = spline[i] + step * direction[i];
    if (!(e - 1e-4 * wolfe * step < spline_energy(trial, t, p_slack, t_slack, p_lambda, t_lambda, pl))) break;
    step *= 0.8;
    t = init_time + step * t_direction;
  }
  for (size_t i = 0; i < n; i++) spline[i] = spline[i] + step * direction[i];
  *piece_time = t;
  free(trial);
}

/* ---- one ADMM iteration: Optimization3D_admm::optimization (Optimization3D_admm.h:29-67) ------------------------------------- */
void port_optimization(double *spline, double *piece_time, double *p_slack, double *t_slack, double *p_lambda, double *t_lambda,
                       double *gnorm_out) {
  Planes pl;
  planes_init(&pl, G.n_tr);
  separate_plane(spline, &pl);
  double *direction = (double *)malloc(sizeof(double) * 3 * (size_t)G.T), t_direction, wolfe, gn;
  descent_direction(spline, *piece_time, p_slack, t_slack, p_lambda, t_lambda, &pl, 0, direction, &t_direction, &wolfe, &gn);
  G.wolfe = wolfe; G.gnorm = gn;
  const double step = position_step(spline, direction);
  line_search(spline, direction, piece_time, t_direction, p_slack, t_slack, p_lambda, t_lambda, &pl, step, wolfe);
  update_slack_lambda(spline, *piece_time, p_slack, t_slack, p_lambda, t_lambda);
  if (gnorm_out) *gnorm_out = gn;
  free(direction);
  planes_free(&pl);
}

/* Optimization3D_multi::update_spline (Optimization3D_multi.h:508-639): the coupled Newton step.  Unknowns: the free control
 * points of every robot (3(T-4) each) + ONE shared piece time; the matrix is block diagonal per robot with a shared arrow
 * row / column (:519-549; the reference factors its sparse view with SimplicialLLT, here dense LLT: same system).  One step
 * for everybody: couple_self_step (:586), the smallest position_step (:589-594), joint Armijo on the summed energies (:605-636). */
static void update_spline_coupled(int u, double *splines, double *piece_time, const double *p_slack, const double *t_slack,
                                  const double *p_lambda, const double *t_lambda, const Planes *pls) {
  const int Pn = G.piece_num, T = G.T, n = 3 * T, ld = n + 1, num = 3 * (T - 4), N = u * num + 1;
  const size_t ns = (size_t)3 * T, np = (size_t)18 * Pn;
  double *Gv = (double *)calloc((size_t)N, sizeof(double)), *H = (double *)calloc((size_t)N * N, sizeof(double));
  double *grad = (double *)malloc(sizeof(double) * (size_t)ld), *hess = (double *)malloc(sizeof(double) * (size_t)ld * ld);
  for (int i = 0; i < u; i++) {
    global_spline_gradient(splines + ns * i, *piece_time, p_slack + np * i, t_slack + (size_t)Pn * i, p_lambda + np * i,
                           t_lambda + (size_t)Pn * i, &pls[i], grad, hess);
    for (int r = 0; r < num; r++) Gv[i * num + r] = grad[6 + r];
    Gv[u * num] += grad[n];
    for (int c = 0; c < num; c++)
      for (int r = 0; r < num; r++) H[(i * num + r) + (size_t)N * (i * num + c)] = hess[(6 + r) + (size_t)ld * (6 + c)];
    for (int r = 0; r < num; r++) {
      H[(i * num + r) + (size_t)N * (u * num)] = hess[(6 + r) + (size_t)ld * n];
      H[(u * num) + (size_t)N * (i * num + r)] = hess[(6 + r) + (size_t)ld * n];
    }
    H[(u * num) + (size_t)N * (u * num)] += hess[n + (size_t)ld * n];
  }
  double *L = (double *)malloc(sizeof(double) * (size_t)N * N), *x = (double *)malloc(sizeof(double) * (size_t)N);
  port_llt(H, L, N);
  for (int i = 0; i < N; i++) x[i] = Gv[i];
  port_llt_solve(L, N, x);
  double wl = 0, gn = 0;
  for (int i = 0; i < N; i++) { x[i] = -x[i]; wl += x[i] * Gv[i]; gn += Gv[i] * Gv[i]; }
  const double wolfe = -wl;
  G.wolfe = wolfe;
  G.gnorm = sqrt(gn) / (double)u;                                               /* :583 */
  double *dirs = (double *)calloc(ns * (size_t)u, sizeof(double));
  for (int i = 0; i < u; i++)
    for (int p = 0; p < T - 4; p++) for (int k = 0; k < 3; k++) dirs[ns * i + (size_t)k * T + 2 + p] = x[i * num + 3 * p + k];
  const double t_direction = x[u * num];
  double *st = (double *)malloc(sizeof(double) * (size_t)u);
  self_steps(splines, dirs, u, 1, st);
  double step = st[0];
  for (int i = 0; i < u; i++) {
    const double s0 = position_step(splines + ns * i, dirs + ns * i);
    if (s0 < step) step = s0;
  }
  if (*piece_time + step * t_direction <= 0) step = -0.95 * *piece_time / t_direction;
  double e0 = 0;
  for (int i = 0; i < u; i++)
    e0 += spline_energy(splines + ns * i, *piece_time, p_slack + np * i, t_slack + (size_t)Pn * i, p_lambda + np * i,
                        t_lambda + (size_t)Pn * i, &pls[i]);
  const double init_time = *piece_time;
  double *trial = (double *)malloc(sizeof(double) * ns * (size_t)u);
  double t = init_time + step * t_direction;
  for (int guard = 0; guard < 2000; guard++) {
    double e1 = 0;
    for (size_t k = 0; k < ns * (size_t)u; k++) trial[k] = splines[k] + step * dirs[k];
    for (int i = 0; i < u; i++)
      e1 += spline_energy(trial + ns * i, t, p_slack + np * i, t_slack + (size_t)Pn * i, p_lambda + np * i,
                          t_lambda + (size_t)Pn * i, &pls[i]);
    if (!(e0 - 1e-4 * wolfe * step < e1)) break;
    step *= 0.8;
    t = init_time + step * t_direction;
  }
  for (size_t k = 0; k < ns * (size_t)u; k++) splines[k] = splines[k] + step * dirs[k];
  *piece_time = t;
  free(Gv); free(H); free(grad); free(hess); free(L); free(x); free(dirs); free(st); free(trial);
}

/* Optimization3D_multi::optimization_decouple (Optimization3D_multi.h:29-118) and, with coupled != 0, ::optimization
 * (:120-174: same plane passes, update_spline, then the slack update of every robot with the shared piece time) */
void port_optimization_multi(int coupled, int u, double *splines, double *piece_time, double *p_slack, double *t_slack,
                             double *p_lambda, double *t_lambda, double *gnorm_out) {
  if (coupled) {
    const int Pc = G.piece_num;
    const size_t nsc = (size_t)3 * G.T, npc = (size_t)18 * Pc;
    Planes *plc = (Planes *)malloc(sizeof(Planes) * (size_t)u);
    for (int i = 0; i < u; i++) { planes_init(&plc[i], G.n_tr); separate_plane_multi(splines + nsc * i, &plc[i]); }
    separate_self(splines, u, plc);
    update_spline_coupled(u, splines, piece_time, p_slack, t_slack, p_lambda, t_lambda, plc);
    for (int i = 0; i < u; i++)
      update_slack_lambda(splines + nsc * i, piece_time[0], p_slack + npc * i, t_slack + (size_t)Pc * i, p_lambda + npc * i,
                          t_lambda + (size_t)Pc * i);
    if (gnorm_out) *gnorm_out = G.gnorm;
    for (int i = 0; i < u; i++) planes_free(&plc[i]);
    free(plc);
    return;
  }
  const int Pn = G.piece_num, T = G.T;
  const size_t ns = (size_t)3 * T, np = (size_t)18 * Pn;
  Planes *pls = (Planes *)malloc(sizeof(Planes) * (size_t)u);
  for (int i = 0; i < u; i++) { planes_init(&pls[i], G.n_tr); separate_plane_multi(splines + ns * i, &pls[i]); }
  separate_self(splines, u, pls);
  double *dirs = (double *)malloc(sizeof(double) * ns * (size_t)u), *tdir = (double *)malloc(sizeof(double) * (size_t)u);
  double *steps = (double *)malloc(sizeof(double) * (size_t)u);
  double gsum = 0, wolfe = 0;
  for (int i = 0; i < u; i++) {
    double gn;
    descent_direction(splines + ns * i, piece_time[i], p_slack + np * i, t_slack + (size_t)Pn * i, p_lambda + np * i,
                      t_lambda + (size_t)Pn * i, &pls[i], 1, dirs + ns * i, &tdir[i], &wolfe, &gn);   /* global `wolfe`: last robot wins (:730) */
    gsum += gn;
  }
  G.wolfe = wolfe;
  G.gnorm = gsum / (double)u;
  self_steps(splines, dirs, u, 0, steps);
  for (int i = 0; i < u; i++) {
    double step = position_step(splines + ns * i, dirs + ns * i);
    if (step < steps[i]) steps[i] = step;
    line_search(splines + ns * i, dirs + ns * i, &piece_time[i], tdir[i], p_slack + np * i, t_slack + (size_t)Pn * i,
                p_lambda + np * i, t_lambda + (size_t)Pn * i, &pls[i], steps[i], wolfe);
  }
  for (int i = 0; i < u; i++)
    update_slack_lambda(splines + ns * i, piece_time[i], p_slack + np * i, t_slack + (size_t)Pn * i, p_lambda + np * i,
                        t_lambda + (size_t)Pn * i);
  if (gnorm_out) *gnorm_out = G.gnorm;
  for (int i = 0; i < u; i++) planes_free(&pls[i]);
  free(pls); free(dirs); free(tdir); free(steps);
}
