// gjk.cuh -- FP64 narrowphase math of the hot path: GJK witness vector (signed-volumes sub-algorithm),
// 49-axis k-DOP separation tests, separating planes, 1-D plane refinement.
//
// Parity contract: every function here reproduces the reference's arithmetic operation by operation
// (same association order, no FMA contraction -- the translation units that include this header are built
// with --fmad=false; the reference is built for plain SSE2, CMakeLists.txt:24).  That makes the discrete
// outputs (k-DOP pass/fail, GJK simplex path, accept/reject of a plane, CCD ladder exponent) bit-identical to
// the reference and (c,d) identical to the last ulp.  Known reference quirks that are part of the behaviour
// and are therefore kept (see comments in place): the witness simplex keeps the *other* edge's vertices in
// the "two edges face the origin" case of the 2-simplex routine, stale barycentric weights are read there,
// and support() keeps the previous support vertex unless a strictly better one exists.
//
// Reference map (lib/opengjk/src/openGJK.c): sv_line = S1D :82-159, sv_tri = S2D :164-394,
// sv_tet = S3D :399-711, support_max = support :714-737, gjk_witness = gjk :754-852.
// HighOrderCCD/CCD/CCD.h: kdop_* = KDOPDCD :354-413, SelfKDOPDCD :535-587, KDOPCCD :416-473,
// SelfKDOPCCD :475-533.  HighOrderCCD/Separate.h: plane_point :18-163, plane_hulls :165-304.
// HighOrderCCD/Optimal_plane.h: refine_d = optimal_d :13-71.
//
// The header is also compiled by g++ (tests/hostsim) so the bit-exactness against the compiled reference can
// be checked on a machine without a GPU; that build is test infrastructure, never a product path.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TOB_HD __host__ __device__ __forceinline__
#ifdef TOB_GJK_INLINE
#define TOB_HDN static __host__ __device__ __forceinline__
#else
#define TOB_HDN static __host__ __device__ __noinline__
#endif
#else
#define TOB_HD inline
#define TOB_HDN inline
#endif

namespace tob {

#define TOB_KDOP_AXES 49

// All indices into the simplex arrays below are compile-time constants after unrolling (run-time positions are handled
// with selects / predicated copies): the simplex then lives in registers instead of local memory, which is what bounded
// the narrowphase kernels.  Operands and operation order are unchanged.
struct Simplex {
  int n;
  double v[4][3];
  int wid[4];
  double lam[4];
};

TOB_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
TOB_HD bool same_sign(double a, double b) { return (a > 0) == (b > 0); }
TOB_HD double nrm2(const double* v) {
  double n2 = 0;
  n2 += v[0] * v[0];
  n2 += v[1] * v[1];
  n2 += v[2] * v[2];
  return n2;
}
TOB_HD double sel3(const double* a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }

// vv = sum_i lam[i] * v[i], accumulated from 0 in index order
TOB_HD void combine(const Simplex& s, int cnt, double* vv) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < cnt) acc += s.lam[i] * s.v[i][j];
    vv[j] = acc;
  }
}

// closest point to the origin on the 1-simplex (v[0]=B, v[1]=A)
TOB_HDN void sv_line(Simplex& s, double* vv) {
  double a[3], b[3], t[3], ft[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    b[i] = s.v[0][i];
    a[i] = s.v[1][i];
    t[i] = b[i] - a[i];
    ft[i] = fabs(t[i]);
  }
  int I = 1;
  if (ft[0] > ft[1]) I = (ft[0] > ft[2]) ? 0 : 2;
  else if (ft[0] < ft[1]) I = (ft[1] > ft[2]) ? 1 : 2;
  else if (ft[0] < ft[2]) I = 2;
  else if (ft[1] < ft[2]) I = 2;

  const double aI = sel3(a, I), bI = sel3(b, I), tI = sel3(t, I);
  double pt = dot3(b, t) / dot3(t, t) * (aI - bI) + bI;
  double det_ap = aI - pt;
  double det_pb = pt - bI;
  bool fa = same_sign(tI, -det_ap);
  bool fb = same_sign(tI, -det_pb);
  if (fa && fb) {
    s.lam[0] = det_ap * -1.0 / tI;
    s.lam[1] = 1 - s.lam[0];
    s.wid[0] = 0; s.wid[1] = 1;
    s.n = 2;
  } else if (!fa) {
    s.lam[0] = 1; s.wid[0] = 0; s.n = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.v[0][i] = s.v[1][i];
  } else {
    s.lam[0] = 1; s.wid[0] = 1; s.n = 1;
  }
  combine(s, s.n, vv);
}

// closest point to the origin on the 2-simplex (v[0]=C, v[1]=B, v[2]=A)
TOB_HDN void sv_tri(Simplex& s, double* vv) {
  double a[3], b[3], c[3], s21[3], s31[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c[i] = s.v[0][i]; b[i] = s.v[1][i]; a[i] = s.v[2][i];
    s21[i] = b[i] - a[i];
    s31[i] = c[i] - a[i];
  }
  // cyclic index pairs (k,l) visited by the reference's "k=l; l=i" walk starting from (1,2): (1,2), (2,0), (0,1)
  double nu[3], fnu[3];
  {
    double m0 = b[1] * c[2] + a[1] * b[2] + c[1] * a[2] - b[1] * a[2] - c[1] * b[2] - a[1] * c[2];
    double m1 = b[2] * c[0] + a[2] * b[0] + c[2] * a[0] - b[2] * a[0] - c[2] * b[0] - a[2] * c[0];
    double m2 = b[0] * c[1] + a[0] * b[1] + c[0] * a[1] - b[0] * a[1] - c[0] * b[1] - a[0] * c[1];
    nu[0] = m0; nu[1] = -m1; nu[2] = m2;      // pow(-1.0,i) * m
    fnu[0] = fabs(nu[0]); fnu[1] = fabs(nu[1]); fnu[2] = fabs(nu[2]);
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
  double nu_max = sel3(nu, I);

  double n[3], nn = 0;
  n[0] = s21[1] * s31[2] - s21[2] * s31[1]; nn += n[0] * n[0];
  n[1] = s21[2] * s31[0] - s21[0] * s31[2]; nn += n[1] * n[1];
  n[2] = s21[0] * s31[1] - s21[1] * s31[0]; nn += n[2] * n[2];
  double inv_len = 1 / sqrt(nn);
#pragma unroll
  for (int i = 0; i < 3; ++i) n[i] = n[i] * inv_len;
  double dna = dot3(n, a);
  double pp0 = dna * sel3(n, J0), pp1 = dna * sel3(n, J1);
  // ss[k] = projection of vertex k (0 = A, 1 = B, 2 = C) on the two kept axes
  const double ss00 = sel3(a, J0), ss01 = sel3(a, J1), ss10 = sel3(b, J0), ss11 = sel3(b, J1), ss20 = sel3(c, J0), ss21 = sel3(c, J1);
  double B[3];
  // (k,l) = (1,2), (2,0), (0,1)
  B[0] = pp0 * ss11 + pp1 * ss20 + ss10 * ss21 - pp0 * ss21 - pp1 * ss10 - ss20 * ss11;
  B[1] = pp0 * ss21 + pp1 * ss00 + ss20 * ss01 - pp0 * ss01 - pp1 * ss20 - ss00 * ss21;
  B[2] = pp0 * ss01 + pp1 * ss10 + ss00 * ss11 - pp0 * ss11 - pp1 * ss00 - ss10 * ss01;
  bool f0 = same_sign(nu_max, B[0]), f1 = same_sign(nu_max, B[1]), f2 = same_sign(nu_max, B[2]);
  double v[3];
  if ((!f1 && !f2) || isnan(n[0])) {
    // both edges through A face the origin: try BA and CA, keep the closer one
    Simplex e;
    e.n = 2; s.n = 2;
    e.lam[0] = 0; e.lam[1] = 0; e.wid[0] = 0; e.wid[1] = 0;
    e.lam[2] = 0; e.lam[3] = 0; e.wid[2] = 0; e.wid[3] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      e.v[0][i] = s.v[1][i];
      e.v[1][i] = s.v[2][i];
      s.v[1][i] = s.v[2][i];
      e.v[2][i] = 0; e.v[3][i] = 0;
    }
    sv_line(e, v);
    sv_line(s, v);
    double vt[3];
    combine(e, e.n, vt);
    combine(s, e.n, v);        // (sic) counted with the other simplex's size: may read a stale weight
    if (dot3(v, v) < dot3(vt, vt)) {
      if (s.n > 1) s.wid[1] = s.wid[1] + 1;      // for (i = 1; i < s.n; ++i) with s.n <= 2
    } else {
      s.n = e.n;               // (sic) weights and labels of BA, vertices of CA are kept
      s.lam[0] = e.lam[0]; s.wid[0] = e.wid[0];
      if (s.n > 1) { s.lam[1] = e.lam[1]; s.wid[1] = e.wid[1]; }
    }
  } else if (f0 && f1 && f2) {
    double inv = 1 / nu_max;
    s.lam[0] = B[2] * inv;
    s.lam[1] = B[1] * inv;
    s.lam[2] = 1 - s.lam[0] - s.lam[1];
    s.wid[0] = 0; s.wid[1] = 1; s.wid[2] = 2;
    s.n = 3;
  } else if (!f2) {            // faces AB
    s.n = 2;
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.v[0][i] = s.v[1][i]; s.v[1][i] = s.v[2][i]; }
    sv_line(s, v);
  } else if (!f1) {            // faces AC
    s.n = 2;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.v[1][i] = s.v[2][i];
    sv_line(s, v);
    if (s.n > 1) s.wid[1] = s.wid[1] + 1;
  } else {                     // faces BC
    s.n = 2;
    sv_line(s, v);
  }
  combine(s, s.n, vv);
}

// vertex of the 3-simplex used by facet candidate `aux` at local slot (2-k): reference TrianglesToTest
TOB_HD int tri_vertex(int aux, int k) {
  // {3,3,3, 1,2,2, 0,0,1}[aux + 3k]
  return (k == 0) ? 3 : (k == 1 ? (aux == 0 ? 1 : 2) : (aux == 2 ? 1 : 0));
}

// s.v[dst] = src (dst is a run-time slot)
TOB_HD void put_vertex(Simplex& s, int dst, const double* src) {
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d == dst) { s.v[d][0] = src[0]; s.v[d][1] = src[1]; s.v[d][2] = src[2]; }
}
TOB_HD void put_wid(Simplex& s, int dst, int val) {
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d == dst) s.wid[d] = val;
}
TOB_HD int get_wid(const Simplex& s, int i) { return i == 0 ? s.wid[0] : (i == 1 ? s.wid[1] : (i == 2 ? s.wid[2] : s.wid[3])); }

// closest point to the origin on the 3-simplex (v[0]=D, v[1]=C, v[2]=B, v[3]=A)
TOB_HDN void sv_tet(Simplex& s, double* vv) {
  double a[3], b[3], c[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { d[i] = s.v[0][i]; c[i] = s.v[1][i]; b[i] = s.v[2][i]; a[i] = s.v[3][i]; }
  double B[4];
  B[0] = -1 * (b[0] * c[1] * d[2] + b[1] * c[2] * d[0] + b[2] * c[0] * d[1] - b[2] * c[1] * d[0] - b[1] * c[0] * d[2] - b[0] * c[2] * d[1]);
  B[1] = +1 * (a[0] * c[1] * d[2] + a[1] * c[2] * d[0] + a[2] * c[0] * d[1] - a[2] * c[1] * d[0] - a[1] * c[0] * d[2] - a[0] * c[2] * d[1]);
  B[2] = -1 * (a[0] * b[1] * d[2] + a[1] * b[2] * d[0] + a[2] * b[0] * d[1] - a[2] * b[1] * d[0] - a[1] * b[0] * d[2] - a[0] * b[2] * d[1]);
  B[3] = +1 * (a[0] * b[1] * c[2] + a[1] * b[2] * c[0] + a[2] * b[0] * c[1] - a[2] * b[1] * c[0] - a[1] * b[0] * c[2] - a[0] * b[2] * c[1]);
  double detM = B[0] + B[1] + B[2] + B[3];

  bool f0 = true, f1 = true, f2 = true, f3 = true;
  const double eps = 1e-13;
  if (fabs(detM) < eps) {
    bool z0 = fabs(B[0]) < eps, z1 = fabs(B[1]) < eps, z2 = fabs(B[2]) < eps, z3 = fabs(B[3]) < eps;
    if (z2 && z3) f1 = false;
    else if (z1 && z3) f2 = false;
    else if (z1 && z2) f3 = false;
    else if (z0 && z3) f1 = false;
    else if (z0 && z2) f1 = false;
    else if (z0 && z1) f2 = false;
    else { f0 = f1 = f2 = f3 = false; }
  } else {
    f0 = same_sign(detM, B[0]); f1 = same_sign(detM, B[1]); f2 = same_sign(detM, B[2]); f3 = same_sign(detM, B[3]);
  }
  int n123 = (int)f1 + (int)f2 + (int)f3;
  double v[3], vt[3];

  if (f0 && n123 == 3) {
    double inv = 1 / detM;
    s.lam[3] = B[0] * inv;
    s.lam[2] = B[1] * inv;
    s.lam[1] = B[2] * inv;
    s.lam[0] = 1 - s.lam[1] - s.lam[2] - s.lam[3];
    s.wid[0] = 0; s.wid[1] = 1; s.wid[2] = 2; s.wid[3] = 3;
    s.n = 4;
  } else if (n123 == 0) {
    // three facets through A face the origin: evaluate ACD, ABD, ABC and keep the closest
    Simplex t;
    t.lam[0] = t.lam[1] = t.lam[2] = t.lam[3] = 0;
    t.wid[0] = t.wid[1] = t.wid[2] = t.wid[3] = 0;
    t.v[3][0] = t.v[3][1] = t.v[3][2] = 0;
    int sid0 = 0, sid1 = 0, sid2 = 0;
    double tl0 = 0, tl1 = 0, tl2 = 0;
    int nclosest = 0;
    double best = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      t.n = 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int vid = tri_vertex(i, k);     // compile-time after unrolling
#pragma unroll
        for (int j = 0; j < 3; ++j) t.v[2 - k][j] = s.v[vid][j];
      }
      sv_tri(t, v);
      combine(t, t.n, vt);
      double dd = dot3(vt, vt);
      if (i == 0 || dd < best) {
        best = dd;
        nclosest = t.n;
        sid0 = tri_vertex(i, t.wid[0]); tl0 = t.lam[0];       // entries >= nclosest are never read
        sid1 = tri_vertex(i, t.wid[1]); tl1 = t.lam[1];
        sid2 = tri_vertex(i, t.wid[2]); tl2 = t.lam[2];
      }
    }
    double keep[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) keep[i][j] = s.v[i][j];
    s.n = nclosest;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (i < nclosest) {
        const int sid = i == 0 ? sid0 : (i == 1 ? sid1 : sid2);
        const double tl = i == 0 ? tl0 : (i == 1 ? tl1 : tl2);
        double kv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) kv[j] = sid == 0 ? keep[0][j] : (sid == 1 ? keep[1][j] : (sid == 2 ? keep[2][j] : keep[3][j]));
        put_vertex(s, nclosest - 1 - i, kv);
        s.lam[i] = tl;
        put_wid(s, nclosest - 1 - i, sid);
      }
    }
  } else if (n123 == 1) {
    // two facets through A face the origin
    Simplex t;
    t.n = 3;
    t.lam[0] = t.lam[1] = t.lam[2] = t.lam[3] = 0;
    t.wid[0] = t.wid[1] = t.wid[2] = t.wid[3] = 0;
    t.v[3][0] = t.v[3][1] = t.v[3][2] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { t.v[0][i] = 0; t.v[1][i] = 0; t.v[2][i] = 0; }
    double best = 0;
    bool used = false;
    int first = 0, second = 0;
    if (!f1) {                 // ACD
#pragma unroll
      for (int i = 0; i < 3; ++i) { t.v[0][i] = s.v[0][i]; t.v[1][i] = s.v[1][i]; t.v[2][i] = s.v[3][i]; }
      sv_tri(t, v);
      combine(t, t.n, vt);
      best = dot3(vt, vt);
      used = true; first = 0;
    }
    if (!f2) {                 // ABD
      if (!used) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { t.v[0][i] = s.v[0][i]; t.v[1][i] = s.v[2][i]; t.v[2][i] = s.v[3][i]; }
        sv_tri(t, v);
        combine(t, t.n, vt);
        best = dot3(vt, vt);
        first = 1;
      } else {
        s.n = 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) { s.v[1][i] = s.v[2][i]; s.v[2][i] = s.v[3][i]; }
        sv_tri(s, v);
        second = 1;
      }
    }
    if (!f3) {                 // ABC
      s.n = 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) { s.v[0][i] = s.v[1][i]; s.v[1][i] = s.v[2][i]; s.v[2][i] = s.v[3][i]; }
      sv_tri(s, v);
      second = 2;
    }
    combine(s, s.n, v);
    if (dot3(v, v) < best) {
      // in place, as the reference: s.wid[s.n-1-i] = tri_vertex(second, s.wid[i]) for i = 0..s.n-1 (later reads see earlier writes)
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < s.n) put_wid(s, s.n - 1 - i, tri_vertex(second, get_wid(s, i)));
    } else {
      s.n = t.n;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < s.n) {
#pragma unroll
          for (int j = 0; j < 3; ++j) s.v[i][j] = t.v[i][j];
          s.lam[i] = t.lam[i];
          put_wid(s, t.n - 1 - i, tri_vertex(first, t.wid[i]));
        }
    }
  } else if (n123 == 2) {
    if (!f1) {                 // ACD
      s.n = 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) s.v[2][i] = s.v[3][i];
      sv_tri(s, v);
    } else if (!f2) {          // ABD
      s.n = 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) { s.v[1][i] = s.v[2][i]; s.v[2][i] = s.v[3][i]; }
      sv_tri(s, v);
      if (s.n > 2) s.wid[2] = s.wid[2] + 1;          // for (i = 2; i < s.n; ++i)
    } else if (!f3) {          // ABC
      s.n = 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) { s.v[0][i] = s.v[1][i]; s.v[1][i] = s.v[2][i]; s.v[2][i] = s.v[3][i]; }
      sv_tri(s, v);
    }
  } else {                     // only BCD faces the origin
    s.n = 3;
    sv_tri(s, v);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < s.n) s.wid[i] = s.wid[i] + 1;
  }
  combine(s, s.n, vv);
}

// support vertex: keeps `cur` unless some vertex is strictly better (scan in index order)
template <int N>
TOB_HD void support_max(const double (*pts)[3], double* cur, const double* dir) {
  double best = dot3(cur, dir);
  double bx = cur[0], by = cur[1], bz = cur[2];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double sv = dot3(pts[i], dir);
    if (sv > best) { best = sv; bx = pts[i][0]; by = pts[i][1]; bz = pts[i][2]; }
  }
  cur[0] = bx; cur[1] = by; cur[2] = bz;
}

// witness vector (closest point of the Minkowski difference A-B to the origin)
template <int NA, int NB>
TOB_HD void gjk_witness(const double (*A)[3], const double (*B)[3], double* vout, unsigned* iters = nullptr) {
  Simplex s;
  s.lam[0] = s.lam[1] = s.lam[2] = s.lam[3] = 0;
  s.wid[0] = s.wid[1] = s.wid[2] = s.wid[3] = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) { s.v[i][0] = 0; s.v[i][1] = 0; s.v[i][2] = 0; }
  double v[3], vm[3], w[3], sa[3], sb[3];
  const double eps_rel2 = 1e-5 * 1e-5;
  const double eps_tot = 1e-15;
  double wmax = 0;
  s.n = 1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = A[0][i] - B[0][i];
    sa[i] = A[0][i];
    sb[i] = B[0][i];
    s.v[0][i] = v[i];
  }
  int k = 0;
  do {
    k++;
    vm[0] = -v[0]; vm[1] = -v[1]; vm[2] = -v[2];
    support_max<NA>(A, sa, vm);
    if (NB > 1) support_max<NB>(B, sb, v);
    w[0] = sa[0] - sb[0]; w[1] = sa[1] - sb[1]; w[2] = sa[2] - sb[2];
    double vv = nrm2(v);
    if ((vv - dot3(v, w)) <= eps_rel2 * vv) break;
    if (vv < eps_rel2) break;
    put_vertex(s, s.n, w);
    s.n++;
    if (s.n == 4) sv_tet(s, v);
    else if (s.n == 3) sv_tri(s, v);
    else if (s.n == 2) sv_line(s, v);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i < s.n) {
        double tn = nrm2(s.v[i]);
        if (tn > wmax) wmax = tn;
      }
    if (nrm2(v) <= (eps_tot * eps_tot * wmax)) break;
  } while ((s.n != 4) && (k != 50));
  vout[0] = v[0]; vout[1] = v[1]; vout[2] = v[2];
  if (iters) *iters += (unsigned)k;   // work counter (bench.py roofline): support + sub-algorithm rounds really executed
}

#if defined(__CUDACC__)
// gjk_witness<6, 1> with the six hull vertices read from memory in every round instead of living in 36 registers for the
// whole loop (gP: 6x3 column-major, the row's control points, L1-resident; q: the point).  Same operands, same operations,
// same order -- the loads are pinned (asm volatile) so that the compiler does not hoist them back out of the loop.  What it
// buys is register room: the narrowphase kernel is bound by the latency of its dependent FP64 chains at 4 warps per scheduler.
__device__ __forceinline__ double ld_pinned(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void gjk_witness_6pt_mem(const double* __restrict__ gP, const double* q, double* vout, unsigned* iters) {
  Simplex s;
  s.lam[0] = s.lam[1] = s.lam[2] = s.lam[3] = 0;
  s.wid[0] = s.wid[1] = s.wid[2] = s.wid[3] = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) { s.v[i][0] = 0; s.v[i][1] = 0; s.v[i][2] = 0; }
  double v[3], vm[3], w[3], sa[3];
  const double eps_rel2 = 1e-5 * 1e-5;
  const double eps_tot = 1e-15;
  double wmax = 0;
  s.n = 1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sa[i] = ld_pinned(gP + 6 * i);
    v[i] = sa[i] - q[i];
    s.v[0][i] = v[i];
  }
  int k = 0;
  do {
    k++;
    vm[0] = -v[0]; vm[1] = -v[1]; vm[2] = -v[2];
    {   // support_max<6>(A, sa, vm)
      double best = dot3(sa, vm);
      double bx = sa[0], by = sa[1], bz = sa[2];
      double px[6], py[6], pz[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { px[i] = ld_pinned(gP + i); py[i] = ld_pinned(gP + 6 + i); pz[i] = ld_pinned(gP + 12 + i); }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double pt[3] = {px[i], py[i], pz[i]};
        double sv = dot3(pt, vm);
        if (sv > best) { best = sv; bx = px[i]; by = py[i]; bz = pz[i]; }
      }
      sa[0] = bx; sa[1] = by; sa[2] = bz;
    }
    w[0] = sa[0] - q[0]; w[1] = sa[1] - q[1]; w[2] = sa[2] - q[2];
    double vv = nrm2(v);
    if ((vv - dot3(v, w)) <= eps_rel2 * vv) break;
    if (vv < eps_rel2) break;
    put_vertex(s, s.n, w);
    s.n++;
    if (s.n == 4) sv_tet(s, v);
    else if (s.n == 3) sv_tri(s, v);
    else if (s.n == 2) sv_line(s, v);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i < s.n) {
        double tn = nrm2(s.v[i]);
        if (tn > wmax) wmax = tn;
      }
    if (nrm2(v) <= (eps_tot * eps_tot * wmax)) break;
  } while ((s.n != 4) && (k != 50));
  vout[0] = v[0]; vout[1] = v[1]; vout[2] = v[2];
  if (iters) *iters += (unsigned)k;
}
#endif

// runtime-sized variants (function-level entry points CCD::GJKDCD with edges / arbitrary vertex counts, CCD.h:17-114)
TOB_HD void support_max_n(const double (*pts)[3], int n, double* cur, const double* dir) {
  double best = dot3(cur, dir);
  int better = -1;
  for (int i = 0; i < n; ++i) {
    double sv = dot3(pts[i], dir);
    if (sv > best) { best = sv; better = i; }
  }
  if (better != -1) { cur[0] = pts[better][0]; cur[1] = pts[better][1]; cur[2] = pts[better][2]; }
}

TOB_HD void gjk_witness_n(const double (*A)[3], int na, const double (*B)[3], int nb, double* vout) {
  Simplex s;
  s.lam[0] = s.lam[1] = s.lam[2] = s.lam[3] = 0;
  s.wid[0] = s.wid[1] = s.wid[2] = s.wid[3] = 0;
  double v[3], vm[3], w[3], sa[3], sb[3];
  const double eps_rel2 = 1e-5 * 1e-5;
  const double eps_tot = 1e-15;
  double wmax = 0;
  s.n = 1;
  for (int i = 0; i < 3; ++i) {
    v[i] = A[0][i] - B[0][i];
    sa[i] = A[0][i];
    sb[i] = B[0][i];
    s.v[0][i] = v[i];
  }
  int k = 0;
  do {
    k++;
    vm[0] = -v[0]; vm[1] = -v[1]; vm[2] = -v[2];
    support_max_n(A, na, sa, vm);
    support_max_n(B, nb, sb, v);
    w[0] = sa[0] - sb[0]; w[1] = sa[1] - sb[1]; w[2] = sa[2] - sb[2];
    double vv = nrm2(v);
    if ((vv - dot3(v, w)) <= eps_rel2 * vv) break;
    if (vv < eps_rel2) break;
    put_vertex(s, s.n, w);
    s.n++;
    if (s.n == 4) sv_tet(s, v);
    else if (s.n == 3) sv_tri(s, v);
    else if (s.n == 2) sv_line(s, v);
    for (int i = 0; i < 4; i++)
      if (i < s.n) {
        double tn = nrm2(s.v[i]);
        if (tn > wmax) wmax = tn;
      }
    if (nrm2(v) <= (eps_tot * eps_tot * wmax)) break;
  } while ((s.n != 4) && (k != 50));
  vout[0] = v[0]; vout[1] = v[1]; vout[2] = v[2];
}

TOB_HD bool kdop_overlap_n(const double (*A)[3], int na, const double (*B)[3], int nb, const double* kdop, double d) {
  for (int k = 0; k < TOB_KDOP_AXES; ++k) {
    double x = kdop[3 * k], y = kdop[3 * k + 1], z = kdop[3 * k + 2];
    double uA = -INFINITY, lA = INFINITY, uB = -INFINITY, lB = INFINITY;
    for (int i = 0; i < na; ++i) {
      double lv = x * A[i][0] + y * A[i][1] + z * A[i][2];
      if (lv < lA) lA = lv;
      if (lv > uA) uA = lv;
    }
    for (int i = 0; i < nb; ++i) {
      double lv = x * B[i][0] + y * B[i][1] + z * B[i][2];
      if (lv < lB) lB = lv;
      if (lv > uB) uB = lv;
    }
    if (uB < lA - d || uA < lB - d) return false;
  }
  return true;
}

// ---- k-DOP -------------------------------------------------------------------------------------------
// level of a point on axis (x,y,z): x*px + y*py + z*pz, left to right
TOB_HD double kdop_level(double x, double y, double z, const double* p) { return x * p[0] + y * p[1] + z * p[2]; }

// extents [lo,hi] of N points on the 49 axes (kdop: 3x49 column-major = axis k at kdop[3k..3k+2])
template <int N>
TOB_HD void kdop_extents(const double (*pts)[3], const double* kdop, double* lo, double* hi) {
  for (int k = 0; k < TOB_KDOP_AXES; ++k) {
    double x = kdop[3 * k], y = kdop[3 * k + 1], z = kdop[3 * k + 2];
    double u = -INFINITY, l = INFINITY;
    for (int i = 0; i < N; ++i) {
      double lv = kdop_level(x, y, z, pts[i]);
      if (lv < l) l = lv;
      if (lv > u) u = lv;
    }
    lo[k] = l; hi[k] = u;
  }
}

// segment extents (precomputed) against one point with gap d
// The 49 axes are tested in 7 groups of 7: inside a group there is no early exit, so the 14 extent loads are issued together
// and the 7 level computations are independent instruction streams (an FP64 result takes ~40 cycles on B200: one axis at
// a time is a ~200-cycle dependent chain per axis, ~10 k cycles for a candidate that passes).  Same comparisons, same
// decision as the axis-by-axis loop of the reference (CCD.h:376-389).
TOB_HD bool kdop_point_overlap(const double* lo, const double* hi, const double* kdop, const double* q, double d,
                               unsigned* groups = nullptr, int axis_begin = 0, int axis_end = TOB_KDOP_AXES) {
  for (int g = axis_begin; g < axis_end; g += 7) {
    bool sep = false;
    if (groups) ++*groups;            // work counter (bench.py roofline): 7-axis groups really evaluated
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int k = g + j;
      const double lv = kdop_level(kdop[3 * k], kdop[3 * k + 1], kdop[3 * k + 2], q);
      const bool below = lv < lo[k] - d, above = hi[k] < lv - d;
      sep = sep | below | above;
    }
    if (sep) return false;
  }
  return true;
}

// ---- single-precision filter of the same gate -----------------------------------------------------------------------
// The gate's outcome is the OR of 98 comparisons (some axis separates), so the order of the axes is free and a comparison
// whose outcome is certain needs no FP64 arithmetic.  Per row, segments.cu stores for every axis k the two thresholds of
//     below_k  <=>  lv < lo_k - d        above_k  <=>  hi_k < lv - d        (lv = level of the point on axis k, FP64)
// relative to the row's centre m, rounded to float:  TL_k = (lo_k - d) - a_k.m,  TH_k = (hi_k + d) - a_k.m.  The point is
// shifted by m in FP64 and rounded to float, its level is three single-precision operations, and a comparison counts as
// decided only when it clears the threshold by more than the allowance
//     E = 2^-21 (|q - m|_1 + max_k |T_k|) + 2^-40 (|m|_1 + d + max_k |T_k|)
// which bounds (with a factor > 1.3 to spare) the sum of: both float conversions of the point and of the axis (2 x 2^-24
// |q-m|_1), the three roundings of the level (3 x 2^-24 |q-m|_1), the float conversion of the threshold and the rounding of
// the difference (2^-24 (2 |T| + |q-m|_1)), and every FP64 rounding of the reference expression and of the threshold
// (< 2^-48 of the magnitudes involved).  Undecided comparisons (and anything non-finite) are re-tested with the reference
// arithmetic, so the decision is the reference's bit for bit; on the benchmark scenes fewer than 1 in 10^4 axes are.
#define TOB_KF_EPS32 4.76837158203125e-07f      // 2^-21
#define TOB_KF_EPS64 9.094947017729282e-13      // 2^-40

struct alignas(8) KfPair { float x, y; };
// thresholds of one row: kf[2k] = TL_k, kf[2k+1] = TH_k, kf[98] = allowance without the point's part (segments.cu)
TOB_HD bool kdop_point_gate(const float* __restrict__ kf, const double* __restrict__ centre, const float* kdop_f,
                            const double* lo, const double* hi, const double* kdop, const double* q, double d,
                            unsigned* groups, unsigned* exact, int axis_begin, int axis_end) {
  const float q0 = (float)(q[0] - centre[0]), q1 = (float)(q[1] - centre[1]), q2 = (float)(q[2] - centre[2]);
  const float E = fmaf(TOB_KF_EPS32, (fabsf(q0) + fabsf(q1)) + fabsf(q2), kf[2 * TOB_KDOP_AXES]);
  const KfPair* __restrict__ th = reinterpret_cast<const KfPair*>(kf);
  for (int g = axis_begin; g < axis_end; g += 7) {
    bool sep = false;
    unsigned unc = 0;
    if (groups) ++*groups;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int k = g + j;
      const float lv = fmaf(kdop_f[3 * k], q0, fmaf(kdop_f[3 * k + 1], q1, kdop_f[3 * k + 2] * q2));
      const KfPair t = th[k];
      const float tl = lv - t.x, thh = lv - t.y;
      sep = sep | (tl < -E) | (thh > E);
      if (!(fabsf(tl) > E) | !(fabsf(thh) > E)) unc |= 1u << j;      // also taken by NaN / inf
    }
    if (sep) return false;
    if (unc) {                      // rare: the reference arithmetic for the axes the filter could not decide
      for (int j = 0; j < 7; ++j)
        if ((unc >> j) & 1u) {
          const int k = g + j;
          if (exact) ++*exact;
          const double lv = kdop_level(kdop[3 * k], kdop[3 * k + 1], kdop[3 * k + 2], q);
          if (lv < lo[k] - d || hi[k] < lv - d) return false;
        }
    }
  }
  return true;
}

// what segments.cu stores for one axis of one row (am = a_k . m in FP64); returns max(|TL|, |TH|)
TOB_HD float kdop_gate_thresholds(double lo, double hi, double d, double am, float* tl, float* th) {
  *tl = (float)((lo - d) - am);
  *th = (float)((hi + d) - am);
  return fmaxf(fabsf(*tl), fabsf(*th));
}
// allowance of a row without the point's part; tmag = max over the axes of kdop_gate_thresholds
TOB_HD float kdop_gate_allowance(float tmag, const double* centre, double d) {
  const double s = (fabs(centre[0]) + fabs(centre[1]) + fabs(centre[2])) + fabs(d) + (double)tmag;
  // rounded up: the product and the sum are inflated by a factor that exceeds their own rounding
  return (TOB_KF_EPS32 * tmag + (float)(TOB_KF_EPS64 * s)) * 1.0000005f + 1e-37f;
}

// two precomputed extent sets with gap d
TOB_HD bool kdop_sets_overlap(const double* loA, const double* hiA, const double* loB, const double* hiB, double d) {
  for (int g = 0; g < TOB_KDOP_AXES; g += 7) {       // groups of 7 axes: 28 loads in flight instead of 4 behind every branch
    bool sep = false;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int k = g + j;
      const bool s0 = hiB[k] < loA[k] - d, s1 = hiA[k] < loB[k] - d;
      sep = sep | s0 | s1;
    }
    if (sep) return false;
  }
  return true;
}

// general (no precomputation) versions used by the function-level entry points
template <int NA, int NB>
TOB_HD bool kdop_overlap(const double (*A)[3], const double (*B)[3], const double* kdop, double d) {
  for (int k = 0; k < TOB_KDOP_AXES; ++k) {
    double x = kdop[3 * k], y = kdop[3 * k + 1], z = kdop[3 * k + 2];
    double uA = -INFINITY, lA = INFINITY, uB = -INFINITY, lB = INFINITY;
    for (int i = 0; i < NA; ++i) {
      double lv = kdop_level(x, y, z, A[i]);
      if (lv < lA) lA = lv;
      if (lv > uA) uA = lv;
    }
    for (int i = 0; i < NB; ++i) {
      double lv = kdop_level(x, y, z, B[i]);
      if (lv < lB) lB = lv;
      if (lv > uB) uB = lv;
    }
    if (uB < lA - d || uA < lB - d) return false;
  }
  return true;
}

// swept vertex set [P + t0*D ; P + t1*D] (CCD.h:419-420 / :119-120)
TOB_HD void swept_points(const double (*P)[3], const double (*D)[3], double t0, double t1, double (*out)[3]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 3; ++j) {
      out[i][j] = P[i][j] + t0 * D[i][j];
      out[i + 6][j] = P[i][j] + t1 * D[i][j];
    }
}

// ---- planes ------------------------------------------------------------------------------------------
// Eigen::Vector3d::norm() = sqrt of the 3-element reduction, associated left to right (checked bitwise
// against the compiled reference)
TOB_HD double eig_norm3(const double* c) { return sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]); }

// segment (6 pts) vs obstacle point: plane n.x + d >= 0 (Separate.h:18-163), in two halves so that a caller can look at the
// distance before it decides (narrow.cu): witness + its norm, then normal and offset.  plane_point() is the two in a row.
TOB_HD double plane_point_witness(const double (*P)[3], const double* q, double* c, unsigned* gjk_iters = nullptr) {
  double Bq[1][3] = {{q[0], q[1], q[2]}};
  gjk_witness<6, 1>(P, Bq, c, gjk_iters);
  return eig_norm3(c);
}
TOB_HD void plane_point_finish(const double* q, double offset, double cn, double* c, double* d) {
  c[0] /= cn; c[1] /= cn; c[2] /= cn;
  double d0 = -c[0] * q[0] - c[1] * q[1] - c[2] * q[2];
  *d = d0 - offset;
}
TOB_HD bool plane_point(const double (*P)[3], const double* q, double distance, double offset, double* c, double* d,
                        unsigned* gjk_iters = nullptr) {
  const double cn = plane_point_witness(P, q, c, gjk_iters);
  if (cn > distance) return false;
  plane_point_finish(q, offset, cn, c, d);
  return true;
}

// segment vs segment (Separate.h:165-304): d = midpoint of the two support levels
TOB_HD bool plane_hulls(const double (*P0)[3], const double (*P1)[3], double distance, double* c, double* d) {
  gjk_witness<6, 6>(P0, P1, c);
  double cn = eig_norm3(c);
  if (cn > distance) return false;
  c[0] /= cn; c[1] /= cn; c[2] /= cn;
  double d0 = INFINITY, d1 = -INFINITY;
  for (int i = 0; i < 6; ++i) {
    // c.dot(row): fixed-size unrolled reduction x0 + (x1 + x2) (checked bitwise against the compiled reference)
    double t = -(c[0] * P1[i][0] + (c[1] * P1[i][1] + c[2] * P1[i][2]));
    if (d0 > t) d0 = t;
  }
  for (int i = 0; i < 6; ++i) {
    double t = -(c[0] * P0[i][0] + (c[1] * P0[i][1] + c[2] * P0[i][2]));
    if (d1 < t) d1 = t;
  }
  *d = 0.5 * (d0 + d1);
  return true;
}

// 1-D Newton on d (Optimal_plane.h:13-71); returns the number of Newton steps taken
TOB_HD int refine_d(const double (*P0)[3], const double (*P1)[3], const double* c, double offset, double margin, double* d_io,
                    int max_iter) {
  double d = *d_io;
  int it = 0;
  while (true) {
    double g = 0, h = 0;
    for (int j = 0; j < 6; ++j) {
      double dist = (P0[j][0] * c[0] + P0[j][1] * c[1] + P0[j][2] * c[2]) + d - 0.5 * offset;
      if (dist < margin) {
        double lg = log(dist / margin);
        double e1 = -(2 * (dist - margin) * lg + (dist - margin) * (dist - margin) / dist);
        double e2 = -(2 * lg + 4 * (dist - margin) / dist - (dist - margin) * (dist - margin) / (dist * dist));
        g += e1; h += e2;
      }
    }
    for (int j = 0; j < 6; ++j) {
      double dist = -(P1[j][0] * c[0] + P1[j][1] * c[1] + P1[j][2] * c[2]) - d - 0.5 * offset;
      if (dist < margin) {
        double lg = log(dist / margin);
        double e1 = -(2 * (dist - margin) * lg + (dist - margin) * (dist - margin) / dist);
        double e2 = -(2 * lg + 4 * (dist - margin) / dist - (dist - margin) * (dist - margin) / (dist * dist));
        g += -e1; h += e2;
      }
    }
    double dir = -g / h;
    d = d + 1.0 * dir;
    it++;
    if (fabs(g) < 1e-2) break;
    if (it >= max_iter) break;
  }
  *d_io = d;
  return it;
}

}  // namespace tob
