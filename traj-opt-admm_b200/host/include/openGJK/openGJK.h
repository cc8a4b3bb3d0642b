// Drop-in for lib/opengjk/include/openGJK/openGJK.h (reference :40-81, the MODIFIED build whose gjk() returns the
// witness vector): same structs and entry point; the arithmetic runs on the device (tob_gjk_batch restates
// lib/opengjk/src/openGJK.c:82-852 operation by operation).
#ifndef OPENGJK_SHADOW_H
#define OPENGJK_SHADOW_H

#include <vector>

#include "trajopt_host.h"

struct bd {
  int numpoints;
  double s[3];
  double** coord;
};

struct simplex {
  int nvrtx;
  double vrtx[4][3];
  int wids[4];
  double lambdas[4];
};

inline double* gjk(struct bd bd1, struct bd bd2, struct simplex* /*s*/) {
  static double c0[3];
  tob_host::Session& S = tob_host::Session::get();
  const int na = bd1.numpoints, nb = bd2.numpoints;
  std::vector<double> A((size_t)3 * na), B((size_t)3 * nb);
  for (int i = 0; i < na; i++) for (int j = 0; j < 3; j++) A[(size_t)j * na + i] = bd1.coord[i][j];
  for (int i = 0; i < nb; i++) for (int j = 0; j < 3; j++) B[(size_t)j * nb + i] = bd2.coord[i][j];
  S.check(tob_gjk_batch(S.ctx(), A.data(), na, B.data(), nb, 1, c0), "tob_gjk_batch");
  return c0;
}

#endif
