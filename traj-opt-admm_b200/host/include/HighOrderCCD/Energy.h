// Drop-in for HighOrderCCD/Energy.h (reference :12-151).  The reference only uses this class for a debug print
// (Optimization3D_multi.h:804-807); SURVEY.md marks it out of scope as a target.  The three functions whose result is a
// plain restatement of an Energy_admm term forward to it (time_weight == 1 in every Main, where they coincide); the
// signatures stay so that callers compile.
#ifndef ENERGY_H
#define ENERGY_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/BVH/BVH.h"
#include "HighOrderCCD/CCD/CCD.h"
#include "HighOrderCCD/Energy_admm.h"

PRJ_BEGIN

class Energy {
 public:
  typedef Eigen::MatrixXd Data;

  // :48-94 -- same sum as Energy_admm::plane_barrier_energy
  static double plane_barrier_energy(const Data& spline, const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                     const std::vector<std::vector<double>>& d_lists) {
    return Energy_admm::plane_barrier_energy(spline, c_lists, d_lists);
  }
  // :96-150 -- Energy_admm::bound_energy with time_weight[sp_id] folded into the piece time (all Mains set it to 1)
  static double bound_energy(const Data& spline, const double& piece_time) {
    for (size_t i = 0; i < time_weight.size(); i++)
      if (time_weight[i] != 1.0) throw std::runtime_error("Energy::bound_energy: only time_weight == 1 is supported");
    return Energy_admm::bound_energy(spline, piece_time);
  }
  // :27-46 -- jerk energy of the whole spline: sum over pieces of the per-piece dynamic term without the time penalty
  static double dynamic_energy(const Data& spline, const double& piece_time) {
    double energy = 0;
    const double kt_keep = kt;
    kt = 0;   // Energy_admm::dynamic_energy adds kt*t^1.1 per piece; this variant (reference :42) has no time term
    try {
      for (int sp_id = 0; sp_id < piece_num; sp_id++) {
        Data bz = spline.block<order_num + 1, 3>(sp_id * (order_num - 2), 0);
        Data c_spline = convert_list[sp_id] * bz;
        energy += Energy_admm::dynamic_energy(c_spline, time_weight[sp_id] * piece_time);
      }
    } catch (...) { kt = kt_keep; throw; }
    kt = kt_keep;
    return energy / ks;   // reference :42 carries no ks factor
  }
  // :17-25
  static double plane_whole_energy(const Data& spline, const double& piece_time, const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                   const std::vector<std::vector<double>>& d_lists) {
    return ks * dynamic_energy(spline, piece_time) + lambda * plane_barrier_energy(spline, c_lists, d_lists) +
           lambda * bound_energy(spline, piece_time) + kt * whole_weight * piece_time;
  }
};

PRJ_END

#endif
