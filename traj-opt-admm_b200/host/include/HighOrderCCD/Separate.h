// Drop-in for HighOrderCCD/Separate.h (reference :13-306): separating plane from the GJK witness vector.
#ifndef SEPARATE_H
#define SEPARATE_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/Optimal_plane.h"
#include "openGJK/openGJK.h"

PRJ_BEGIN

class Separate {
 public:
  typedef Eigen::MatrixXd Data;

  // :18-163 segment (6 control points) vs obstacle point: c = v/|v|, d = -c.q - offset; false when |v| > distance
  static bool opengjk(const Data& position, const Data& _position, const double& distance, Eigen::Vector3d& c, double& d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    if (position.rows() != 6 || _position.rows() != 1) throw std::runtime_error("Separate::opengjk: expects 6x3 and 1x3");
    Data P = position, q = _position;
    uint8_t ok = 0; double cc[3] = {0, 0, 0}, dd = 0;
    S.check(tob_plane_point_batch(S.ctx(), P.data(), q.data(), 1, distance, &ok, cc, &dd), "tob_plane_point_batch");
    c = Eigen::Vector3d(cc[0], cc[1], cc[2]);   // the reference leaves the raw witness in c on rejection
    if (ok) d = dd;
    return ok != 0;
  }

  // :165-304 segment vs segment: d = midpoint of the two support levels
  static bool selfgjk(const Data& position, const Data& _position, const double& distance, Eigen::Vector3d& c, double& d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    if (position.rows() != 6 || _position.rows() != 6) throw std::runtime_error("Separate::selfgjk: expects two 6x3 polygons");
    Data P0 = position, P1 = _position;
    uint8_t ok = 0; double cc[3] = {0, 0, 0}, dd = 0;
    S.check(tob_plane_hulls_batch(S.ctx(), P0.data(), P1.data(), 1, distance, 0, &ok, cc, &dd), "tob_plane_hulls_batch");
    c = Eigen::Vector3d(cc[0], cc[1], cc[2]);   // the reference leaves the raw witness in c on rejection
    if (ok) d = dd;
    return ok != 0;
  }
};

PRJ_END

#endif
