// Drop-in for HighOrderCCD/Optimal_plane.h (reference :8-778): per-pair plane refinement.
//   optimal_d       :13-71    1-D Newton on d (always used for inter-robot planes)            -> tob_refine_d_batch
//   optimal_cd      :160-293  Newton on the tangent angles of c, (sub-segment, obstacle point)  -> tob_optimal_cd_batch
//   self_optimal_cd :620-773  (theta, phi, d) Newton, inter-robot plane                         -> tob_self_optimal_cd_batch
// The last two belong to the persistent-plane mode ("optimal_plane":1); inside optimization() the device keeps the live
// planes itself (include/trajopt_b200.h, tob_planes_reset), these statics serve callers that refine single planes.
#ifndef OPTIMAL_PLANE_H
#define OPTIMAL_PLANE_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "trajopt_host.h"

PRJ_BEGIN

class Optimal_plane {
 public:
  typedef Eigen::MatrixXd Data;

  static void optimal_d(const Data& position, const Data& _position, const Eigen::Vector3d& c, double& d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data P0 = position, P1 = _position;
    double cc[3] = {c(0), c(1), c(2)};
    S.check(tob_refine_d_batch(S.ctx(), P0.data(), P1.data(), cc, 1, &d), "tob_refine_d_batch");
  }

  static void optimal_cd(const Data& position, const Data& _position, Eigen::Vector3d& c, double& d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data P = position;
    double q[3] = {_position(0, 0), _position(0, 1), _position(0, 2)}, cc[3] = {c(0), c(1), c(2)};
    S.check(tob_optimal_cd_batch(S.ctx(), P.data(), q, 1, cc, &d, nullptr), "tob_optimal_cd_batch");
    c = Eigen::Vector3d(cc[0], cc[1], cc[2]);
  }

  static void self_optimal_cd(const Data& position, const Data& _position, Eigen::Vector3d& c, double& d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data P0 = position, P1 = _position;
    double cc[3] = {c(0), c(1), c(2)};
    S.check(tob_self_optimal_cd_batch(S.ctx(), P0.data(), P1.data(), 1, cc, &d, nullptr), "tob_self_optimal_cd_batch");
    c = Eigen::Vector3d(cc[0], cc[1], cc[2]);
  }
};

PRJ_END

#endif
