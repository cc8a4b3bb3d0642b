// Drop-in for HighOrderCCD/Step.h (reference :12-416): CCD conservative step bounds.  plane_step (:258-310) and
// mix_step (:313-411) are dead code in the reference and omitted.
#ifndef STEP_H
#define STEP_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/BVH/BVH.h"
#include "HighOrderCCD/CCD/CCD.h"

PRJ_BEGIN

class Step {
 public:
  typedef Eigen::MatrixXd Data;
  typedef std::pair<unsigned int, unsigned int> id_pair;

  // :21-110  largest 0.8^k so that hull(P u P+step*D) stays offset away from every cloud point
  static double position_step(const Data& spline, const Data& direction, const std::vector<Eigen::RowVector3d>& /*vertex_list*/, BVH& /*bvh*/) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    double step = 1.0;
    S.check(tob_position_step(S.ctx(), spline.data(), direction.data(), &step), "tob_position_step");
    return step;
  }

  // :112-182  one shared step for all robots
  static void couple_self_step(const std::vector<Data>& spline_list, const std::vector<Data>& direction_list, double& step, BVH& /*bvh*/) {
    if (step != 1.0) throw std::runtime_error("Step::couple_self_step: the device ladder starts at step = 1 (the reference's only call site, "
                                              "Optimization3D_multi.h:586)");
    std::vector<double> out(1, 1.0);
    run(spline_list, direction_list, 1, out);
    step = out[0];
  }

  // :184-256  one step per robot, pairs resolved in (time slot, pair) order
  static void self_step(const std::vector<Data>& spline_list, const std::vector<Data>& direction_list, std::vector<double>& step_list, BVH& /*bvh*/) {
    step_list.assign(spline_list.size(), 1.0);
    run(spline_list, direction_list, 0, step_list);
  }

 private:
  static void run(const std::vector<Data>& sl, const std::vector<Data>& dl, int coupled, std::vector<double>& out) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    const int u = (int)sl.size();
    if (u != uav_num || (int)dl.size() != u) throw std::runtime_error("Step::self_step: needs uav_num splines and directions");
    const size_t n = (size_t)3 * sl[0].rows();
    std::vector<double> s(n * u), d(n * u);
    for (int i = 0; i < u; i++) {
      std::memcpy(&s[n * i], sl[i].data(), n * sizeof(double));
      std::memcpy(&d[n * i], dl[i].data(), n * sizeof(double));
    }
    S.check(tob_self_step(S.ctx(), s.data(), d.data(), u, coupled, out.data()), "tob_self_step");
  }
};

PRJ_END

#endif
