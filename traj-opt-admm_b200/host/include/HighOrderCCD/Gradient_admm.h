// Drop-in for HighOrderCCD/Gradient_admm.h (reference :8-676): gradient and Hessian blocks of the augmented Lagrangian.
// Live entry points keep their signatures; the dead ones of the reference (spline_gradient :166-236,
// plane_barrier_gradient :238-278, bound_gradient :281-329, target_gradient :624-631 -- never called) are omitted.
#ifndef GRADIENT_ADMM_H
#define GRADIENT_ADMM_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "trajopt_host.h"

PRJ_BEGIN

class Gradient_admm {
 public:
  typedef Eigen::MatrixXd Data;

  // :13-65  dense (3T+1) gradient and (3T+1)^2 Hessian from the PSD-projected piece blocks
  static void global_spline_gradient(const Data& spline, const double& piece_time, const Data& p_slack, const Eigen::VectorXd& t_slack,
                                     const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                     const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                     const std::vector<std::vector<double>>& d_lists, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    tob_host::set_planes(c_lists, d_lists);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    const int n = 3 * (int)spline.rows();
    grad.resize(n + 1);
    hessian.resize(n + 1, n + 1);
    S.check(tob_global_gradient(S.ctx(), 0, &v.st, grad.data(), hessian.data()), "tob_global_gradient");
  }

  // :67-164  one piece, before the PSD projection: 19-vector and 19x19 block
  static void local_spline_gradient(const Data& spline, const double& piece_time, const Data& p_slack, const Eigen::VectorXd& t_slack,
                                    const Data& p_lambda, const Eigen::VectorXd& t_lambda,
                                    const std::vector<std::vector<Eigen::Vector3d>>& c_lists,
                                    const std::vector<std::vector<double>>& d_lists, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian,
                                    int sp_id) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    tob_host::check_state_shapes(spline, p_slack, t_slack, p_lambda, t_lambda);
    if (sp_id < 0 || sp_id >= piece_num) throw std::runtime_error("Gradient_admm::local_spline_gradient: sp_id out of range");
    tob_host::set_planes(c_lists, d_lists);
    tob_host::StateView v(spline, piece_time, p_slack, t_slack, p_lambda, t_lambda);
    std::vector<double> g((size_t)19 * piece_num), h((size_t)361 * piece_num);
    S.check(tob_piece_blocks(S.ctx(), 0, &v.st, 0, g.data(), h.data()), "tob_piece_blocks");
    grad = Eigen::Map<Eigen::VectorXd>(g.data() + 19 * sp_id, 19);
    hessian = Eigen::Map<Eigen::MatrixXd>(h.data() + 361 * sp_id, 19, 19);
  }

  // :331-407  one sub-segment: plane barrier terms only, 18-vector and 18x18
  static void local_plane_barrier_gradient(int tr_id, const Data& spline, const std::vector<Eigen::Vector3d>& c_list,
                                           const std::vector<double>& d_list, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    std::vector<std::vector<Eigen::Vector3d>> cl(piece_num * res);
    std::vector<std::vector<double>> dl(piece_num * res);
    cl.at(tr_id) = c_list; dl.at(tr_id) = d_list;
    tob_host::set_planes(cl, dl);
    grad.resize(18); hessian.resize(18, 18);
    S.check(tob_row_blocks(S.ctx(), spline.data(), 1.0, tr_id, 0, grad.data(), hessian.data(), nullptr, nullptr, nullptr), "tob_row_blocks");
  }

  // :409-572  one sub-segment: velocity / acceleration bound terms incl. the time derivatives
  static void local_bound_gradient(int tr_id, const Data& spline, const double& piece_time, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian,
                                   double& g_t, double& h_t, Eigen::VectorXd& partgrad) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    grad.resize(18); hessian.resize(18, 18); partgrad.resize(18);
    S.check(tob_row_blocks(S.ctx(), spline.data(), piece_time, tr_id, 1, grad.data(), hessian.data(), &g_t, &h_t, partgrad.data()), "tob_row_blocks");
  }

  // :574-622  one piece of the slack problem
  static void slack_gradient(const Data& c_spline, const double& piece_time, const Data& p_part, const double& t_part, const Data& p_lambda,
                             const double& t_lambda, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data cs = c_spline, pp = p_part, pl = p_lambda;
    grad.resize(19); hessian.resize(19, 19);
    S.check(tob_slack_terms(S.ctx(), cs.data(), piece_time, pp.data(), t_part, pl.data(), t_lambda, 1, nullptr, grad.data(), hessian.data()),
            "tob_slack_terms");
  }

  // :633-671
  static void dynamic_gradient(const Data& p_part, const double& t_part, Eigen::VectorXd& grad, Eigen::MatrixXd& hessian, double& g_t,
                               double& h_t, Eigen::VectorXd& partgrad) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync();
    Data pp = p_part;
    Eigen::VectorXd g(19);
    Eigen::MatrixXd h(19, 19);
    S.check(tob_slack_terms(S.ctx(), nullptr, 0.0, pp.data(), t_part, nullptr, 0.0, 0, nullptr, g.data(), h.data()), "tob_slack_terms");
    grad = g.head(18);
    hessian = h.block(0, 0, 18, 18);
    g_t = g(18);
    h_t = h(18, 18);
    partgrad = h.block(0, 18, 18, 1);
  }
};

PRJ_END

#endif
