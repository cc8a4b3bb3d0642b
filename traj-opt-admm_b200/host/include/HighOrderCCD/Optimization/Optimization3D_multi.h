// Drop-in for HighOrderCCD/Optimization/Optimization3D_multi.h (reference :21-815): multi-UAV ADMM iteration.
//   optimization_decouple()   :29-118   per-robot Newton systems, per-robot steps      -> tob_optimization(mode 0)
//   optimization()            :120-174  one coupled system with a shared piece time    -> tob_optimization(mode 1)
// plus the function-level entry points with the reference's signatures.
#ifndef OPTIMIZATION3D_MULTI_H
#define OPTIMIZATION3D_MULTI_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/Energy.h"
#include "HighOrderCCD/Energy_admm.h"
#include "HighOrderCCD/Gradient_admm.h"
#include "HighOrderCCD/Step.h"
#include "HighOrderCCD/Separate.h"
#include "HighOrderCCD/Optimization/Optimization3D_admm.h"

#include <vector>
#include <ctime>

PRJ_BEGIN

class Optimization3D_multi {
 public:
  typedef Eigen::MatrixXd Data;
  typedef Eigen::SparseMatrix<double> SpMat;

  static void optimization_decouple(std::vector<Data>& spline_list, std::vector<double>& piece_time_list, std::vector<Data>& p_slack_list,
                                    std::vector<Eigen::VectorXd>& t_slack_list, std::vector<Data>& p_lambda_list,
                                    std::vector<Eigen::VectorXd>& t_lambda_list, const std::vector<Eigen::RowVector3d>& /*vertex_list*/,
                                    BVH& /*bvh*/) {
    std::vector<tob_state> st(uav_num);
    if ((int)piece_time_list.size() != uav_num) throw std::runtime_error("optimization_decouple: piece_time_list must have uav_num entries");
    for (int i = 0; i < uav_num; i++) st[i].piece_time = &piece_time_list[i];
    run(st, 0, spline_list, p_slack_list, t_slack_list, p_lambda_list, t_lambda_list);
  }

  static void optimization(std::vector<Data>& spline_list, double& piece_time, std::vector<Data>& p_slack_list,
                           std::vector<Eigen::VectorXd>& t_slack_list, std::vector<Data>& p_lambda_list,
                           std::vector<Eigen::VectorXd>& t_lambda_list, const std::vector<Eigen::RowVector3d>& /*vertex_list*/, BVH& /*bvh*/) {
    std::vector<tob_state> st(uav_num);
    for (int i = 0; i < uav_num; i++) st[i].piece_time = &piece_time;   // one shared piece time
    run(st, 1, spline_list, p_slack_list, t_slack_list, p_lambda_list, t_lambda_list);
  }

  // :176-235 -- same pipeline as the single-UAV one
  static void separate_plane(const Data& spline, const std::vector<Eigen::RowVector3d>& vertex_list,
                             std::vector<std::vector<Eigen::Vector3d>>& c_list, std::vector<std::vector<double>>& d_list, BVH& bvh) {
    Optimization3D_admm::separate_plane(spline, vertex_list, c_list, d_list, bvh);
  }

  // :237-342 -- inter-robot planes of every time slot, APPENDED to the per-robot lists
  static void separate_self(const std::vector<Data>& spline_list, std::vector<std::vector<std::vector<Eigen::Vector3d>>>& self_c_lists,
                            std::vector<std::vector<std::vector<double>>>& self_d_lists, BVH& /*bvh*/) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    const int u = (int)spline_list.size(), n_tr = piece_num * res;
    if (u != uav_num) throw std::runtime_error("separate_self: needs uav_num splines");
    const size_t n = (size_t)3 * spline_list[0].rows();
    std::vector<double> s(n * u);
    for (int i = 0; i < u; i++) std::memcpy(&s[n * i], spline_list[i].data(), n * sizeof(double));
    std::vector<uint32_t> off((size_t)u * n_tr + 1);
    std::vector<double> c(3 << 14), d(1 << 14);
    uint64_t total = 0;
    for (int pass = 0; pass < 2; pass++) {
      S.check(tob_separate_self(S.ctx(), s.data(), u, off.data(), c.data(), d.data(), d.size(), &total), "tob_separate_self");
      if (total <= d.size()) break;
      c.resize(3 * total); d.resize(total);
    }
    self_c_lists.resize(u); self_d_lists.resize(u);
    for (int i = 0; i < u; i++) {
      self_c_lists[i].resize(n_tr); self_d_lists[i].resize(n_tr);
      tob_host::csr_to_planes(off.data() + (size_t)i * n_tr, c.data(), d.data(), n_tr, self_c_lists[i], self_d_lists[i], true);
    }
  }

  // :344-506 -- identical to the single-UAV update
  static void update_slack_lambda(const Data& spline, const double& piece_time, Data& p_slack, Eigen::VectorXd& t_slack, Data& p_lambda,
                                  Eigen::VectorXd& t_lambda) {
    Optimization3D_admm::update_slack_lambda(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
  }

  // :641-657
  static double spline_energy(const std::vector<Data>& spline_list, const double& piece_time, const std::vector<Data>& p_slack_list,
                              const std::vector<Eigen::VectorXd>& t_slack_list, const std::vector<Data>& p_lambda_list,
                              const std::vector<Eigen::VectorXd>& t_lambda_list,
                              const std::vector<std::vector<std::vector<Eigen::Vector3d>>>& c_lists,
                              const std::vector<std::vector<std::vector<double>>>& d_lists) {
    double e = 0;
    for (int i = 0; i < uav_num; i++)
      e += Energy_admm::spline_energy(spline_list[i], piece_time, p_slack_list[i], t_slack_list[i], p_lambda_list[i], t_lambda_list[i],
                                      c_lists[i], d_lists[i]);
    return e;
  }

  // :659-752 -- dense LLT variant; accumulates gnorm, overwrites wolfe
  static void spline_descent_direction(const Data& spline, Data& direction, const double& piece_time, double& t_direction, const Data& p_slack,
                                       const Eigen::VectorXd& t_slack, const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                       const std::vector<std::vector<Eigen::Vector3d>>& c_list,
                                       const std::vector<std::vector<double>>& d_list) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_list, d_list);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    direction.resize(spline.rows(), 3);
    double w = 0, gn = 0;
    S.check(tob_descent_direction(S.ctx(), 0, &v.st, 1, direction.data(), &t_direction, &w, &gn), "tob_descent_direction");
    wolfe = w;
    gnorm += gn;
  }

  // :754-811 -- Armijo backtracking from the given step bound
  static void spline_line_search(Data& spline, const Data& direction, double& piece_time, const double& t_direction, const Data& p_slack,
                                 const Eigen::VectorXd& t_slack, const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                 const std::vector<std::vector<Eigen::Vector3d>>& c_list, const std::vector<std::vector<double>>& d_list,
                                 double& step) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_list, d_list);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    S.check(tob_line_search(S.ctx(), 0, &v.st, direction.data(), t_direction, wolfe, &step), "tob_line_search");
    piece_time = v.pt;
  }

 private:
  static void run(std::vector<tob_state>& st, int mode, std::vector<Data>& spline_list, std::vector<Data>& p_slack_list,
                  std::vector<Eigen::VectorXd>& t_slack_list, std::vector<Data>& p_lambda_list, std::vector<Eigen::VectorXd>& t_lambda_list) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    if ((int)spline_list.size() != uav_num) throw std::runtime_error("Optimization3D_multi: needs uav_num splines");
    for (int i = 0; i < uav_num; i++) {
      tob_host::check_state_shapes(spline_list[i], p_slack_list[i], t_slack_list[i], p_lambda_list[i], t_lambda_list[i]);
      st[i].spline = spline_list[i].data(); st[i].p_slack = p_slack_list[i].data(); st[i].t_slack = t_slack_list[i].data();
      st[i].p_lambda = p_lambda_list[i].data(); st[i].t_lambda = t_lambda_list[i].data();
    }
    gnorm = S.optimization(st.data(), uav_num, mode);   // all GPUs of the session (TRAJOPT_B200_GPUS), robots sharded
    double w = 0;
    if (tob_last_wolfe(S.ctx(), &w) == 0) wolfe = w;
  }
};

PRJ_END

#endif
