// Drop-in for HighOrderCCD/Optimization/Optimization3D_admm.h (reference :21-583): single-UAV ADMM iteration.
//   optimization()            one fused, device-resident iteration (tob_optimization): planes never leave the GPU
//   separate_plane() ...      function-level entry points with the reference's signatures (parity tests, partial adoption)
// Globals written like the reference does: gnorm, wolfe (Optimization3D_admm.h:477,499).
#ifndef OPTIMIZATION3D_ADMM_H
#define OPTIMIZATION3D_ADMM_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/CCD/CCD.h"
#include "HighOrderCCD/Energy.h"
#include "HighOrderCCD/Energy_admm.h"
#include "HighOrderCCD/Gradient_admm.h"
#include "HighOrderCCD/Step.h"
#include "HighOrderCCD/Separate.h"

PRJ_BEGIN

class Optimization3D_admm {
 public:
  typedef Eigen::MatrixXd Data;
  typedef Eigen::SparseMatrix<double> SpMat;

  // :29-67
  static void optimization(Data& spline, double& piece_time, Data& p_slack, Eigen::VectorXd& t_slack, Data& p_lambda,
                           Eigen::VectorXd& t_lambda, const std::vector<Eigen::RowVector3d>& /*vertex_list*/, BVH& /*bvh*/) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_state st;
    st.spline = spline.data(); st.piece_time = &piece_time; st.p_slack = p_slack.data(); st.t_slack = t_slack.data();
    st.p_lambda = p_lambda.data(); st.t_lambda = t_lambda.data();
    double gn = 0;
    S.check(tob_optimization(S.ctx(), &st, 1, 0, &gn), "tob_optimization");
    gnorm = gn;
    double w = 0;
    if (tob_last_wolfe(S.ctx(), &w) == 0) wolfe = w;
  }

  // :69-197 (optimal_plane = 0)
  static void separate_plane(const Data& spline, const std::vector<Eigen::RowVector3d>& /*vertex_list*/,
                             std::vector<std::vector<Eigen::Vector3d>>& c_lists, std::vector<std::vector<double>>& d_lists, BVH& /*bvh*/) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    const int n_tr = piece_num * res;
    std::vector<uint32_t> off(n_tr + 1);
    std::vector<double> c(3 << 16), d(1 << 16);
    uint64_t total = 0;
    for (int pass = 0; pass < 2; pass++) {
      S.check(tob_separate_planes(S.ctx(), spline.data(), 1, 0, off.data(), c.data(), d.data(), d.size(), &total), "tob_separate_planes");
      if (total <= d.size()) break;
      c.resize(3 * total); d.resize(total);
    }
    tob_host::csr_to_planes(off.data(), c.data(), d.data(), n_tr, c_lists, d_lists, false);
  }

  // :199-229
  static void update_spline(Data& spline, double& piece_time, const Data& p_slack, const Eigen::VectorXd& t_slack, const Data& p_lambda,
                            const Eigen::VectorXd& t_lambda, const std::vector<Eigen::RowVector3d>& vertex_list, BVH& bvh,
                            const std::vector<std::vector<Eigen::Vector3d>>& c_lists, const std::vector<std::vector<double>>& d_lists) {
    Data direction;
    double t_direction;
    spline_descent_direction(spline, direction, piece_time, t_direction, p_slack, t_slack, p_lambda, t_lambda, c_lists, d_lists);
    spline_line_search(spline, direction, piece_time, t_direction, p_slack, t_slack, p_lambda, t_lambda, vertex_list, bvh, c_lists, d_lists);
  }

  // :231-398
  static void update_slack_lambda(const Data& spline, const double& piece_time, Data& p_slack, Eigen::VectorXd& t_slack, Data& p_lambda,
                                  Eigen::VectorXd& t_lambda) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    double pt = piece_time;
    tob_state st;
    st.spline = const_cast<double*>(spline.data()); st.piece_time = &pt; st.p_slack = p_slack.data(); st.t_slack = t_slack.data();
    st.p_lambda = p_lambda.data(); st.t_lambda = t_lambda.data();
    S.check(tob_update_slack_lambda(S.ctx(), &st), "tob_update_slack_lambda");
  }

  // :400-503
  static int spline_descent_direction(const Data& spline, Data& direction, const double& piece_time, double& t_direction, const Data& p_slack,
                                      const Eigen::VectorXd& t_slack, const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                      const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                      const std::vector<std::vector<double>>& d_lists) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_lists, d_lists);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    direction.resize(spline.rows(), 3);
    double w = 0, gn = 0;
    S.check(tob_descent_direction(S.ctx(), 0, &v.st, 0, direction.data(), &t_direction, &w, &gn), "tob_descent_direction");
    wolfe = w;
    gnorm = gn;
    return 1;
  }

  // :505-557  CCD step bound, clamp on the piece time, Armijo backtracking with factor 0.8
  static void spline_line_search(Data& spline, const Data& direction, double& piece_time, const double& t_direction, const Data& p_slack,
                                 const Eigen::VectorXd& t_slack, const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                 const std::vector<Eigen::RowVector3d>& /*vertex_list*/, BVH& /*bvh*/,
                                 const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                 const std::vector<std::vector<double>>& d_lists) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_lists, d_lists);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    double step = -1.0;                         // < 0: the step bound is Step::position_step, computed on the device
    S.check(tob_line_search(S.ctx(), 0, &v.st, direction.data(), t_direction, wolfe, &step), "tob_line_search");
    piece_time = v.pt;                          // spline was updated in place through the state view
  }

  // :559-578  diagnostic energy (barrier + bounds + per-piece jerk/time energy)
  static double spline_energy(const Data& spline, const double& piece_time, const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                              const std::vector<std::vector<double>>& d_lists) {
    double energy = lambda * Energy_admm::plane_barrier_energy(spline, c_lists, d_lists) + lambda * Energy_admm::bound_energy(spline, piece_time);
    for (int sp_id = 0; sp_id < piece_num; sp_id++) {
      Data c_spline = convert_list[sp_id] * spline.block<order_num + 1, 3>(sp_id * (order_num - 2), 0);
      energy += Energy_admm::dynamic_energy(c_spline, piece_time);
    }
    return energy;
  }
};

PRJ_END

#endif
