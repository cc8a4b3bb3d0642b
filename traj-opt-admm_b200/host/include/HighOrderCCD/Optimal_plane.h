// Drop-in for HighOrderCCD/Optimal_plane.h (reference :8-778): per-pair plane refinement.
//   optimal_d       :13-71    1-D Newton on d (always used for inter-robot planes)            -> tob_refine_d_batch
// optimal_cd (:160-293) and self_optimal_cd (:620-773) belong to the persistent-plane mode ("optimal_plane":1), which
// Config File/3D.json switches off; tob_set_params rejects that mode, so they are not provided here.
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
};

PRJ_END

#endif
