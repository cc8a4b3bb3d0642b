// Drop-in for HighOrderCCD/Energy_admm.h (reference :8-219): augmented-Lagrangian objective.  Every function marshals
// its Eigen arguments into the C ABI; the ragged plane lists become the resident CSR plane set of the context.
#ifndef ENERGY_ADMM_H
#define ENERGY_ADMM_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "trajopt_host.h"

PRJ_BEGIN

class Energy_admm {
 public:
  typedef Eigen::MatrixXd Data;

  // :16-44
  static double spline_energy(const Data& spline, const double& piece_time, const Data& p_slack, const Eigen::VectorXd& t_slack,
                              const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                              const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                              const std::vector<std::vector<double>>& d_lists) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_lists, d_lists);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    double e = 0;
    S.check(tob_spline_energy(S.ctx(), 0, &v.st, &e), "tob_spline_energy");
    return e;
  }

  // :46-96
  static double plane_barrier_energy(const Data& spline, const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                     const std::vector<std::vector<double>>& d_lists) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::set_planes(c_lists, d_lists);
    double e = 0;
    S.check(tob_plane_barrier_energy(S.ctx(), 0, spline.data(), &e), "tob_plane_barrier_energy");
    return e;
  }

  // :98-170
  static double bound_energy(const Data& spline, const double& piece_time) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    double e = 0;
    S.check(tob_bound_energy(S.ctx(), spline.data(), piece_time, &e), "tob_bound_energy");
    return e;
  }

  // :172-190 (one piece)
  static double slack_energy(const Data& c_spline, const double& piece_time, const Data& p_part, const double& t_part,
                             const Data& p_lambda, const double& t_lambda) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data cs = c_spline, pp = p_part, pl = p_lambda;
    double e = 0;
    S.check(tob_slack_terms(S.ctx(), cs.data(), piece_time, pp.data(), t_part, pl.data(), t_lambda, 1, &e, nullptr, nullptr), "tob_slack_terms");
    return e;
  }

  // :192-197 (unused by the optimisers; trivial closed form kept for source compatibility)
  static double target_energy(const Eigen::Vector3d& endpoint, const Eigen::Vector3d& target) { return 0.5 * (endpoint - target).squaredNorm(); }

  // :199-215 (one piece)
  static double dynamic_energy(const Data& p_part, const double& t_part) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data pp = p_part;
    double e = 0;
    S.check(tob_slack_terms(S.ctx(), nullptr, 0.0, pp.data(), t_part, nullptr, 0.0, 0, &e, nullptr, nullptr), "tob_slack_terms");
    return e;
  }
};

PRJ_END

#endif
