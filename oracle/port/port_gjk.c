/* port_gjk.c -- GJK witness vector, plain-C restatement of the reference's vendored and modified openGJK
 * (lib/opengjk/src/openGJK.c): S1D :82-159 (sv_line), S2D :164-394 (sv_tri), S3D :399-711 (sv_tet), support :714-737,
 * gjk :754-852 (non-ADAPTIVEFP build, lib/opengjk/CMakeLists.txt:37-46; returns the witness vector v, :844-849).
 *
 * TEST INFRASTRUCTURE ONLY (see port.h).  The reference's behaviour is kept operation by operation, including its quirks:
 * support() keeps the previous support vertex unless a strictly better one exists (:722-735); in the "two edges face the
 * origin" case of S2D the barycentric weights of the other sub-simplex are read (:350-372); the out-of-bounds read of
 * indexJ on exact ties (:222) is pinned to index 0 (it only matters for degenerate triangles, where the isnan guard decides).
 */
#include <math.h>

#include "port.h"

#define isnan_(x) isnan(x)

typedef struct Simplex {
  int n;
  double v[4][3];
  int wid[4];
  double lam[4];
} Simplex;

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline int same_sign(double a, double b) { return (a > 0) == (b > 0); }
static inline double nrm2(const double* v) {
  double n2 = 0;
  n2 += v[0] * v[0];
  n2 += v[1] * v[1];
  n2 += v[2] * v[2];
  return n2;
}

// vv = sum_i lam[i] * v[i], accumulated from 0 in index order
static inline void combine(const Simplex *s, int cnt, double* vv) {
  for (int j = 0; j < 3; ++j) {
    double acc = 0;
    for (int i = 0; i < cnt; ++i) acc += s->lam[i] * s->v[i][j];
    vv[j] = acc;
  }
}

// closest point to the origin on the 1-simplex (v[0]=B, v[1]=A)
static void sv_line(Simplex *s, double* vv) {
  double a[3], b[3], t[3], ft[3];
  for (int i = 0; i < 3; ++i) {
    b[i] = s->v[0][i];
    a[i] = s->v[1][i];
    t[i] = b[i] - a[i];
    ft[i] = fabs(t[i]);
  }
  int I = 1;
  if (ft[0] > ft[1]) I = (ft[0] > ft[2]) ? 0 : 2;
  else if (ft[0] < ft[1]) I = (ft[1] > ft[2]) ? 1 : 2;
  else if (ft[0] < ft[2]) I = 2;
  else if (ft[1] < ft[2]) I = 2;

  double pt = dot3(b, t) / dot3(t, t) * (a[I] - b[I]) + b[I];
  double det_ap = a[I] - pt;
  double det_pb = pt - b[I];
  int fa = same_sign(t[I], -det_ap);
  int fb = same_sign(t[I], -det_pb);
  if (fa && fb) {
    s->lam[0] = det_ap * -1.0 / t[I];
    s->lam[1] = 1 - s->lam[0];
    s->wid[0] = 0; s->wid[1] = 1;
    s->n = 2;
  } else if (!fa) {
    s->lam[0] = 1; s->wid[0] = 0; s->n = 1;
    for (int i = 0; i < 3; ++i) s->v[0][i] = s->v[1][i];
  } else {
    s->lam[0] = 1; s->wid[0] = 1; s->n = 1;
  }
  combine(s, s->n, vv);
}

// closest point to the origin on the 2-simplex (v[0]=C, v[1]=B, v[2]=A)
static void sv_tri(Simplex *s, double* vv) {
  double a[3], b[3], c[3], s21[3], s31[3];
  for (int i = 0; i < 3; ++i) {
    c[i] = s->v[0][i]; b[i] = s->v[1][i]; a[i] = s->v[2][i];
    s21[i] = b[i] - a[i];
    s31[i] = c[i] - a[i];
  }
  // cyclic index pairs (k,l) visited by the reference's "k=l; l=i" walk starting from (1,2)
  const int K[3] = {1, 2, 0}, L[3] = {2, 0, 1};
  double nu[3], fnu[3];
  for (int i = 0; i < 3; ++i) {
    int k = K[i], l = L[i];
    double m = b[k] * c[l] + a[k] * b[l] + c[k] * a[l] - b[k] * a[l] - c[k] * b[l] - a[k] * c[l];
    nu[i] = (i == 1) ? -m : m;   // pow(-1.0,i) * m
    fnu[i] = fabs(nu[i]);
  }
  // the reference initialises indexJ[2] = {-1} i.e. {-1,0}; the no-branch-taken case (exact ties) reads out of
  // bounds there (undefined); we pin it to J = {0,0}, which only matters when the triangle is degenerate and the
  // isnan() guard below takes over anyway.
  int I = 1, J0 = 0, J1 = 0;
  if (fnu[0] > fnu[1]) {
    if (fnu[0] > fnu[2]) { I = 0; J0 = 1; J1 = 2; } else { J0 = 0; J1 = 1; I = 2; }
  } else if (fnu[0] < fnu[1]) {
    if (fnu[1] > fnu[2]) { J0 = 0; I = 1; J1 = 2; } else { J0 = 0; J1 = 1; I = 2; }
  } else if (fnu[0] < fnu[2]) { J0 = 0; J1 = 1; I = 2; }
  double nu_max = nu[I];

  double n[3], nn = 0;
  for (int i = 0; i < 3; ++i) {
    int k = K[i], l = L[i];
    n[i] = s21[k] * s31[l] - s21[l] * s31[k];
    nn += n[i] * n[i];
  }
  double inv_len = 1 / sqrt(nn);
  for (int i = 0; i < 3; ++i) n[i] = n[i] * inv_len;
  double dna = dot3(n, a);
  double pp0 = dna * n[J0], pp1 = dna * n[J1];
  double ss[3][2] = {{a[J0], a[J1]}, {b[J0], b[J1]}, {c[J0], c[J1]}};
  double B[3];
  for (int i = 0; i < 3; ++i) {
    int k = K[i], l = L[i];
    B[i] = pp0 * ss[k][1] + pp1 * ss[l][0] + ss[k][0] * ss[l][1] - pp0 * ss[l][1] - pp1 * ss[k][0] - ss[l][0] * ss[k][1];
  }
  int f0 = same_sign(nu_max, B[0]), f1 = same_sign(nu_max, B[1]), f2 = same_sign(nu_max, B[2]);
  double v[3];
  if ((!f1 && !f2) || isnan_(n[0])) {
    // both edges through A face the origin: try BA and CA, keep the closer one
    Simplex e;
    e.n = 2; s->n = 2;
    e.lam[0] = 0; e.lam[1] = 0; e.wid[0] = 0; e.wid[1] = 0;
    for (int i = 0; i < 3; ++i) {
      e.v[0][i] = s->v[1][i];
      e.v[1][i] = s->v[2][i];
      s->v[1][i] = s->v[2][i];
    }
    sv_line(&e, v);
    sv_line(s, v);
    double vt[3];
    combine(&e, e.n, vt);
    combine(s, e.n, v);        // (sic) counted with the other simplex's size: may read a stale weight
    if (dot3(v, v) < dot3(vt, vt)) {
      for (int i = 1; i < s->n; ++i) s->wid[i] = s->wid[i] + 1;
    } else {
      s->n = e.n;               // (sic) weights and labels of BA, vertices of CA are kept
      for (int i = 0; i < s->n; ++i) { s->lam[i] = e.lam[i]; s->wid[i] = e.wid[i]; }
    }
  } else if (f0 && f1 && f2) {
    double inv = 1 / nu_max;
    s->lam[0] = B[2] * inv;
    s->lam[1] = B[1] * inv;
    s->lam[2] = 1 - s->lam[0] - s->lam[1];
    s->wid[0] = 0; s->wid[1] = 1; s->wid[2] = 2;
    s->n = 3;
  } else if (!f2) {            // faces AB
    s->n = 2;
    for (int i = 0; i < 3; ++i) { s->v[0][i] = s->v[1][i]; s->v[1][i] = s->v[2][i]; }
    sv_line(s, v);
  } else if (!f1) {            // faces AC
    s->n = 2;
    for (int i = 0; i < 3; ++i) s->v[1][i] = s->v[2][i];
    sv_line(s, v);
    for (int i = 1; i < s->n; ++i) s->wid[i] = s->wid[i] + 1;
  } else {                     // faces BC
    s->n = 2;
    sv_line(s, v);
  }
  combine(s, s->n, vv);
}

// vertex of the 3-simplex used by facet candidate `aux` at local slot (2-k): reference TrianglesToTest
static inline int tri_vertex(int aux, int k) {
  // {3,3,3, 1,2,2, 0,0,1}[aux + 3k]
  return (k == 0) ? 3 : (k == 1 ? (aux == 0 ? 1 : 2) : (aux == 2 ? 1 : 0));
}

// closest point to the origin on the 3-simplex (v[0]=D, v[1]=C, v[2]=B, v[3]=A)
static void sv_tet(Simplex *s, double* vv) {
  double a[3], b[3], c[3], d[3];
  for (int i = 0; i < 3; ++i) { d[i] = s->v[0][i]; c[i] = s->v[1][i]; b[i] = s->v[2][i]; a[i] = s->v[3][i]; }
  double B[4];
  B[0] = -1 * (b[0] * c[1] * d[2] + b[1] * c[2] * d[0] + b[2] * c[0] * d[1] - b[2] * c[1] * d[0] - b[1] * c[0] * d[2] - b[0] * c[2] * d[1]);
  B[1] = +1 * (a[0] * c[1] * d[2] + a[1] * c[2] * d[0] + a[2] * c[0] * d[1] - a[2] * c[1] * d[0] - a[1] * c[0] * d[2] - a[0] * c[2] * d[1]);
  B[2] = -1 * (a[0] * b[1] * d[2] + a[1] * b[2] * d[0] + a[2] * b[0] * d[1] - a[2] * b[1] * d[0] - a[1] * b[0] * d[2] - a[0] * b[2] * d[1]);
  B[3] = +1 * (a[0] * b[1] * c[2] + a[1] * b[2] * c[0] + a[2] * b[0] * c[1] - a[2] * b[1] * c[0] - a[1] * b[0] * c[2] - a[0] * b[2] * c[1]);
  double detM = B[0] + B[1] + B[2] + B[3];

  int f[4] = {1, 1, 1, 1};
  const double eps = 1e-13;
  if (fabs(detM) < eps) {
    int z0 = fabs(B[0]) < eps, z1 = fabs(B[1]) < eps, z2 = fabs(B[2]) < eps, z3 = fabs(B[3]) < eps;
    if (z2 && z3) f[1] = 0;
    else if (z1 && z3) f[2] = 0;
    else if (z1 && z2) f[3] = 0;
    else if (z0 && z3) f[1] = 0;
    else if (z0 && z2) f[1] = 0;
    else if (z0 && z1) f[2] = 0;
    else { f[0] = f[1] = f[2] = f[3] = 0; }
  } else {
    for (int i = 0; i < 4; ++i) f[i] = same_sign(detM, B[i]);
  }
  int n123 = (int)f[1] + (int)f[2] + (int)f[3];
  double v[3], vt[3];

  if (f[0] && n123 == 3) {
    double inv = 1 / detM;
    s->lam[3] = B[0] * inv;
    s->lam[2] = B[1] * inv;
    s->lam[1] = B[2] * inv;
    s->lam[0] = 1 - s->lam[1] - s->lam[2] - s->lam[3];
    s->wid[0] = 0; s->wid[1] = 1; s->wid[2] = 2; s->wid[3] = 3;
    s->n = 4;
  } else if (n123 == 0) {
    // three facets through A face the origin: evaluate ACD, ABD, ABC and keep the closest
    Simplex t;
    t.lam[0] = t.lam[1] = t.lam[2] = t.lam[3] = 0;
    t.wid[0] = t.wid[1] = t.wid[2] = t.wid[3] = 0;
    int sid[4] = {0, 0, 0, 0};
    double tl[4] = {0, 0, 0, 0};
    int nclosest = 0;
    double best = 0;
    for (int i = 0; i < 3; ++i) {
      t.n = 3;
      for (int k = 0; k < 3; ++k) {
        int vid = tri_vertex(i, k);
        for (int j = 0; j < 3; ++j) t.v[2 - k][j] = s->v[vid][j];
      }
      sv_tri(&t, v);
      combine(&t, t.n, vt);
      double dd = dot3(vt, vt);
      if (i == 0 || dd < best) {
        best = dd;
        nclosest = t.n;
        for (int l = 0; l < nclosest; ++l) { sid[l] = tri_vertex(i, t.wid[l]); tl[l] = t.lam[l]; }
      }
    }
    double keep[4][3];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) keep[i][j] = s->v[i][j];
    s->n = nclosest;
    for (int i = 0; i < s->n; ++i) {
      for (int j = 0; j < 3; ++j) s->v[nclosest - 1 - i][j] = keep[sid[i]][j];
      s->lam[i] = tl[i];
      s->wid[nclosest - 1 - i] = sid[i];
    }
  } else if (n123 == 1) {
    // two facets through A face the origin
    Simplex t;
    t.n = 3;
    t.lam[0] = t.lam[1] = t.lam[2] = t.lam[3] = 0;
    t.wid[0] = t.wid[1] = t.wid[2] = t.wid[3] = 0;
    double best = 0;
    int used = 0;
    int first = 0, second = 0;
    if (!f[1]) {               // ACD
      for (int i = 0; i < 3; ++i) { t.v[0][i] = s->v[0][i]; t.v[1][i] = s->v[1][i]; t.v[2][i] = s->v[3][i]; }
      sv_tri(&t, v);
      combine(&t, t.n, vt);
      best = dot3(vt, vt);
      used = 1; first = 0;
    }
    if (!f[2]) {               // ABD
      if (!used) {
        for (int i = 0; i < 3; ++i) { t.v[0][i] = s->v[0][i]; t.v[1][i] = s->v[2][i]; t.v[2][i] = s->v[3][i]; }
        sv_tri(&t, v);
        combine(&t, t.n, vt);
        best = dot3(vt, vt);
        first = 1;
      } else {
        s->n = 3;
        for (int i = 0; i < 3; ++i) { s->v[1][i] = s->v[2][i]; s->v[2][i] = s->v[3][i]; }
        sv_tri(s, v);
        second = 1;
      }
    }
    if (!f[3]) {               // ABC
      s->n = 3;
      for (int i = 0; i < 3; ++i) { s->v[0][i] = s->v[1][i]; s->v[1][i] = s->v[2][i]; s->v[2][i] = s->v[3][i]; }
      sv_tri(s, v);
      second = 2;
    }
    combine(s, s->n, v);
    if (dot3(v, v) < best) {
      for (int i = 0; i < s->n; ++i) s->wid[s->n - 1 - i] = tri_vertex(second, s->wid[i]);   // in place, as the reference
    } else {
      s->n = t.n;
      for (int i = 0; i < s->n; ++i) {
        for (int j = 0; j < 3; ++j) s->v[i][j] = t.v[i][j];
        s->lam[i] = t.lam[i];
        s->wid[t.n - 1 - i] = tri_vertex(first, t.wid[i]);
      }
    }
  } else if (n123 == 2) {
    if (!f[1]) {               // ACD
      s->n = 3;
      for (int i = 0; i < 3; ++i) s->v[2][i] = s->v[3][i];
      sv_tri(s, v);
    } else if (!f[2]) {        // ABD
      s->n = 3;
      for (int i = 0; i < 3; ++i) { s->v[1][i] = s->v[2][i]; s->v[2][i] = s->v[3][i]; }
      sv_tri(s, v);
      for (int i = 2; i < s->n; ++i) s->wid[i] = s->wid[i] + 1;
    } else if (!f[3]) {        // ABC
      s->n = 3;
      for (int i = 0; i < 3; ++i) { s->v[0][i] = s->v[1][i]; s->v[1][i] = s->v[2][i]; s->v[2][i] = s->v[3][i]; }
      sv_tri(s, v);
    }
  } else {                     // only BCD faces the origin
    s->n = 3;
    sv_tri(s, v);
    for (int i = 0; i < s->n; ++i) s->wid[i] = s->wid[i] + 1;
  }
  combine(s, s->n, vv);
}


/* support (:714-737): scan in index order, keep `cur` unless strictly better */
static void support_max(const double (*pts)[3], int n, double *cur, const double *dir) {
  double best = dot3(cur, dir);
  int better = -1;
  for (int i = 0; i < n; ++i) {
    double sv = dot3(pts[i], dir);
    if (sv > best) { best = sv; better = i; }
  }
  if (better != -1) { cur[0] = pts[better][0]; cur[1] = pts[better][1]; cur[2] = pts[better][2]; }
}

/* gjk (:754-852): witness vector of the minimum distance between hull(A) and hull(B) */
void port_gjk_witness(const double (*A)[3], int na, const double (*B)[3], int nb, double *vout) {
  Simplex s;
  s.lam[0] = s.lam[1] = s.lam[2] = s.lam[3] = 0;
  s.wid[0] = s.wid[1] = s.wid[2] = s.wid[3] = 0;
  double v[3], vm[3], w[3], sa[3], sb[3];
  const double eps_rel2 = 1e-5 * 1e-5;   /* eps_rel22 (:763) */
  const double eps_tot = 1e-15;          /* eps_tot22 = eps_tot * eps_tot (:762) */
  double wmax = 0;
  s.n = 1;
  for (int i = 0; i < 3; ++i) {           /* :778-786: start from the first vertices */
    v[i] = A[0][i] - B[0][i];
    sa[i] = A[0][i];
    sb[i] = B[0][i];
    s.v[0][i] = v[i];
  }
  int k = 0;
  do {
    k++;
    vm[0] = -v[0]; vm[1] = -v[1]; vm[2] = -v[2];
    support_max(A, na, sa, vm);
    support_max(B, nb, sb, v);
    w[0] = sa[0] - sb[0]; w[1] = sa[1] - sb[1]; w[2] = sa[2] - sb[2];
    double vv = nrm2(v);
    if ((vv - dot3(v, w)) <= eps_rel2 * vv) break;          /* :804 */
    if (vv < eps_rel2) break;                                /* :810 */
    int i = s.n;
    s.v[i][0] = w[0]; s.v[i][1] = w[1]; s.v[i][2] = w[2];
    s.n++;
    if (s.n == 4) sv_tet(&s, v);
    else if (s.n == 3) sv_tri(&s, v);
    else if (s.n == 2) sv_line(&s, v);
    for (i = 0; i < s.n; i++) {                              /* :829-835 */
      double tn = nrm2(s.v[i]);
      if (tn > wmax) wmax = tn;
    }
    if (nrm2(v) <= (eps_tot * eps_tot * wmax)) break;        /* :837 */
  } while ((s.n != 4) && (k != 50));                         /* :841 */
  vout[0] = v[0]; vout[1] = v[1]; vout[2] = v[2];
}
