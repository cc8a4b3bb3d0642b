// Drop-in for HighOrderCCD/CCD/CCD.h (reference :12-589): the seven static predicates keep their signatures; each call
// evaluates ONE pair through the batched device primitives (tob_gjk_batch / tob_kdop_batch).  The iteration itself
// (Optimization3D_*::optimization) never goes through these per-pair calls -- it runs the fused device pipeline.
#ifndef CCD_H
#define CCD_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "trajopt_host.h"

PRJ_BEGIN

class CCD {
 public:
  typedef Eigen::MatrixXd Data;

  // :17-114  gjk(position, _position); collide iff |v|^2 <= d^2
  static bool GJKDCD(const Data& position, const Data& _position, const double& d) {
    return gjk_dist2(position, _position) <= d * d;
  }
  // :116-225  hull of [P+tMin*D ; P+tMax*D] against the point set
  static bool GJKCCD(const Data& position, const Data& direction, const Data& _position, const double& d, const double& tMin,
                     const double& tMax) {
    return gjk_dist2(swept(position, direction, tMin, tMax), _position) <= d * d;
  }
  // :227-352
  static bool SelfGJKCCD(const Data& position, const Data& direction, const Data& _position, const Data& _direction, const double& d,
                         const double& tMin, const double& tMax, const double& _tMin, const double& _tMax) {
    return gjk_dist2(swept(position, direction, tMin, tMax), swept(_position, _direction, _tMin, _tMax)) <= d * d;
  }
  // :354-413
  static bool KDOPDCD(const Data& position, const Data& _position, const double& d) { return kdop(position, _position, d); }
  // :416-473
  static bool KDOPCCD(const Data& position, const Data& direction, const Data& _position, const double& d, const double& tMin,
                      const double& tMax) {
    return kdop(swept(position, direction, tMin, tMax), _position, d);
  }
  // :475-533
  static bool SelfKDOPCCD(const Data& position, const Data& direction, const Data& _position, const Data& _direction, const double& d,
                          const double& tMin, const double& tMax, const double& _tMin, const double& _tMax) {
    return kdop(swept(position, direction, tMin, tMax), swept(_position, _direction, _tMin, _tMax), d);
  }
  // :535-587
  static bool SelfKDOPDCD(const Data& position, const Data& _position, const double& d) { return kdop(position, _position, d); }

 private:
  // [P + t0*D ; P + t1*D]  (reference :119-120, :419-420); built without FMA contraction (see host/Makefile)
  static Data swept(const Data& P, const Data& D, double t0, double t1) {
    Data A(2 * P.rows(), 3);
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < P.rows(); i++) {
        A(i, j) = P(i, j) + t0 * D(i, j);
        A(i + P.rows(), j) = P(i, j) + t1 * D(i, j);
      }
    return A;
  }
  static double gjk_dist2(const Data& A, const Data& B) {
    tob_host::Session& S = tob_host::Session::get();
    double v[3];
    Data a = A, b = B;
    S.check(tob_gjk_batch(S.ctx(), a.data(), (int)a.rows(), b.data(), (int)b.rows(), 1, v), "tob_gjk_batch");
    return v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  }
  static bool kdop(const Data& A, const Data& B, double d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    uint8_t f = 0;
    Data a = A, b = B;
    S.check(tob_kdop_batch(S.ctx(), a.data(), (int)a.rows(), b.data(), (int)b.rows(), 1, d, &f), "tob_kdop_batch");
    return f != 0;
  }
};

PRJ_END

#endif
