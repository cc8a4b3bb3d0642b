// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" driver over the UNMODIFIED reference sources.  Nothing from
// /root/reference is copied into this repository: this file only #includes the
// reference headers where they lie (see oracle/Makefile for the -I paths and
// the list of reference translation units compiled beside it) and forwards
// flat C arrays to the reference's own static entry points.  The resulting
// oracle/_ref/libtrajopt_ref.so is the parity oracle and the "reference" CPU
// baseline of bench.py.
//
// The only logic restated here is the one-time set-up that the reference keeps
// inside its executables (Main/admmPathPlanning3D.cpp:403-414 k-DOP / AABB axis
// matrices, :448-468 time weights / combination / Conversion, :294-338
// subdivision tables) because a file with main() cannot be linked.
//
// Layout conventions (same as include/trajopt_b200.h): every matrix is
// column-major FP64 exactly as Eigen::MatrixXd stores it.

#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include "HighOrderCCD/Optimization/Optimization3D_admm.h"
#include "HighOrderCCD/Optimization/Optimization3D_multi.h"

USE_PRJ_NAMESPACE

typedef Eigen::MatrixXd Data;

namespace {

BVH* g_bvh = nullptr;
std::vector<Eigen::RowVector3d> g_vertex_list;
std::ostringstream g_sink;
std::streambuf* g_cout_buf = nullptr;

struct Quiet {  // the reference prints on every call; swallow it
  Quiet() { g_sink.str(""); g_cout_buf = std::cout.rdbuf(g_sink.rdbuf()); }
  ~Quiet() { std::cout.rdbuf(g_cout_buf); }
};

Data map_mat(const double* p, int rows, int cols) {
  return Eigen::Map<const Data>(p, rows, cols);
}

void planes_from_csr(const unsigned* off, const double* c, const double* d, int n_tr,
                     std::vector<std::vector<Eigen::Vector3d>>& cl, std::vector<std::vector<double>>& dl) {
  cl.assign(n_tr, std::vector<Eigen::Vector3d>());
  dl.assign(n_tr, std::vector<double>());
  for (int t = 0; t < n_tr; t++)
    for (unsigned k = off[t]; k < off[t + 1]; k++) {
      cl[t].push_back(Eigen::Vector3d(c[3 * k], c[3 * k + 1], c[3 * k + 2]));
      dl[t].push_back(d[k]);
    }
}

long planes_to_csr(const std::vector<std::vector<Eigen::Vector3d>>& cl, const std::vector<std::vector<double>>& dl,
                   unsigned* off, double* c, double* d, long cap) {
  long n = 0;
  off[0] = 0;
  for (size_t t = 0; t < cl.size(); t++) {
    for (size_t k = 0; k < cl[t].size(); k++) {
      if (n < cap) {
        c[3 * n] = cl[t][k](0); c[3 * n + 1] = cl[t][k](1); c[3 * n + 2] = cl[t][k](2);
        d[n] = dl[t][k];
      }
      n++;
    }
    off[t + 1] = (unsigned)n;
  }
  return n;
}

long lists_to_csr(const std::vector<std::vector<unsigned int>>& lists, unsigned* off, unsigned* ids, long cap) {
  long n = 0;
  off[0] = 0;
  for (size_t t = 0; t < lists.size(); t++) {
    for (size_t k = 0; k < lists[t].size(); k++) {
      if (n < cap) ids[n] = lists[t][k];
      n++;
    }
    off[t + 1] = (unsigned)n;
  }
  return n;
}

}  // namespace

extern "C" {

// ---- set-up -----------------------------------------------------------------
// Mirrors Main/admmPathPlanning3D.cpp:368-414,444-468 and init_variable :294-338.
void ref_setup(int piece_num_, int res_, int uav_num_, double lambda_, double margin_, double offset_,
               double mu_, double vel_limit_, double acc_limit_, double ks_, double kt_, int optimal_plane_) {
  Quiet q;
  static bool axes_done = false;
  piece_num = piece_num_; res = res_; uav_num = uav_num_;
  lambda = lambda_; margin = margin_; offset = offset_; mu = mu_;
  vel_limit = vel_limit_; acc_limit = acc_limit_; ks = ks_; kt = kt_;
  epsilon = 0.1; is_optimal_plane = optimal_plane_; automove = true; iter = 0; gnorm = 1;
  int dim = kdop_axis.size();
  kdop_matrix.resize(3, dim);
  for (int k = 0; k < dim; k++) {
    if (!axes_done) kdop_axis[k].normalize();
    kdop_matrix.col(k) = kdop_axis[k];
  }
  axes_done = true;
  aabb_matrix.resize(3, 3);
  for (int k = 0; k < 3; k++) aabb_matrix.col(k) = aabb_axis[k];

  time_weight.assign(piece_num, 1.0);
  whole_weight = piece_num;
  trajectory_num = (order_num + 1) + (piece_num - 1) * (order_num + 1 - 3);
  combination = Combination<40>::value();
  Conversion<order_num>::convert_matrix();
  M_dynamic = Dynamic3D<order_num, der_num>::dynamic_matrix();

  subdivide_tree.clear(); subdivide_tree.resize(piece_num * res);
  A_list.assign(piece_num * res, std::vector<Eigen::MatrixXd>());
  A_vel_list.assign(piece_num * res, std::vector<Eigen::MatrixXd>());
  A_acc_list.assign(piece_num * res, std::vector<Eigen::MatrixXd>());
  Eigen::MatrixXd basis, tmp_basis;
  Eigen::Matrix3d I; I.setIdentity();
  for (int k = 0; k < res; k++) {
    double a = k / double(res), b = (k + 1) / double(res);
    Blossom<order_num>::coefficient(basis, a, b);
    for (int i = 0; i < piece_num; i++) {
      std::pair<double, double> range(a, b);
      subdivide_tree[i * res + k] = std::make_tuple(i, range, basis * convert_list[i]);
      tmp_basis = basis * convert_list[i];
      A_list[i * res + k].resize(order_num + 1);
      A_vel_list[i * res + k].resize(order_num);
      A_acc_list[i * res + k].resize(order_num - 1);
      for (int j = 0; j <= order_num; j++) {
        Eigen::MatrixXd A = Eigen::kroneckerProduct(tmp_basis.row(j), I);
        A.transposeInPlace();
        A_list[i * res + k][j] = A;
        if (j < order_num) {
          A = Eigen::kroneckerProduct(tmp_basis.row(j + 1), I) - Eigen::kroneckerProduct(tmp_basis.row(j), I);
          A_vel_list[i * res + k][j] = A;
        }
        if (j < order_num - 1) {
          A = Eigen::kroneckerProduct(tmp_basis.row(j + 2), I) - 2 * Eigen::kroneckerProduct(tmp_basis.row(j + 1), I) +
              Eigen::kroneckerProduct(tmp_basis.row(j), I);
          A_acc_list[i * res + k][j] = A;
        }
      }
    }
  }
}

int ref_trajectory_num() { return trajectory_num; }

// basis: n_tr x 36 (each 6x6 col-major), weight: n_tr (b-a), convert: piece_num x 36,
// mdyn: 36, kdop: 3 x 49 col-major
void ref_get_tables(double* basis, double* weight, double* convert, double* mdyn, double* kdop) {
  for (size_t t = 0; t < subdivide_tree.size(); t++) {
    std::memcpy(basis + 36 * t, std::get<2>(subdivide_tree[t]).data(), 36 * sizeof(double));
    weight[t] = std::get<1>(subdivide_tree[t]).second - std::get<1>(subdivide_tree[t]).first;
  }
  for (int i = 0; i < piece_num; i++) std::memcpy(convert + 36 * i, convert_list[i].data(), 36 * sizeof(double));
  std::memcpy(mdyn, M_dynamic.data(), 36 * sizeof(double));
  std::memcpy(kdop, kdop_matrix.data(), 3 * kdop_axis.size() * sizeof(double));
}

// BVH::InitPointcloud (BVH.cpp:53-92) + the vertex_list copy of Main :435-440.
void ref_init_pointcloud(const double* V, int n) {
  Quiet q;
  delete g_bvh;
  g_bvh = new BVH();
  Data Vm = map_mat(V, n, 3);
  g_bvh->InitPointcloud(Vm);
  g_vertex_list.resize(n);
  for (int i = 0; i < n; i++) g_vertex_list[i] = Vm.row(i);
}

// ---- broadphase -------------------------------------------------------------
long ref_dcd_collision(const double* spline, double d, unsigned* off, unsigned* ids, long cap) {
  std::vector<std::vector<unsigned int>> pairs;
  g_bvh->DCDCollision(map_mat(spline, trajectory_num, 3), pairs, d);
  return lists_to_csr(pairs, off, ids, cap);
}

long ref_ccd_collision(const double* spline, const double* direction, double d, unsigned* off, unsigned* ids, long cap) {
  std::vector<std::vector<unsigned int>> pairs;
  g_bvh->CCDCollision(map_mat(spline, trajectory_num, 3), map_mat(direction, trajectory_num, 3), pairs, d);
  return lists_to_csr(pairs, off, ids, cap);
}

// P, D: u matrices 6x3 col-major back to back. pairs: 2 ints each.
long ref_self_dcd(const double* P, int u, double d, unsigned* pairs, long cap) {
  std::vector<Data> Pl;
  for (int i = 0; i < u; i++) Pl.push_back(map_mat(P + 18 * i, 6, 3));
  std::vector<std::pair<unsigned, unsigned>> out;
  BVH b;
  b.SelfDCDCollision(Pl, out, d);
  for (size_t i = 0; i < out.size() && (long)i < cap; i++) { pairs[2 * i] = out[i].first; pairs[2 * i + 1] = out[i].second; }
  return out.size();
}

long ref_self_ccd(const double* P, const double* D, int u, double d, unsigned* pairs, long cap) {
  std::vector<Data> Pl, Dl;
  for (int i = 0; i < u; i++) { Pl.push_back(map_mat(P + 18 * i, 6, 3)); Dl.push_back(map_mat(D + 18 * i, 6, 3)); }
  std::vector<std::pair<unsigned, unsigned>> out;
  BVH b;
  b.SelfCCDCollision(Pl, Dl, out, d);
  for (size_t i = 0; i < out.size() && (long)i < cap; i++) { pairs[2 * i] = out[i].first; pairs[2 * i + 1] = out[i].second; }
  return out.size();
}

// ---- narrowphase primitives ------------------------------------------------
// sub-segment control points P_tr = basis_tr * bz exactly as the reference forms them
void ref_segment_points(const double* spline, int tr_id, double* P) {
  int sp_id = std::get<0>(subdivide_tree[tr_id]);
  Eigen::MatrixXd basis = std::get<2>(subdivide_tree[tr_id]);
  Data sp = map_mat(spline, trajectory_num, 3);
  Eigen::MatrixXd bz;
  bz = sp.block<order_num + 1, 3>(sp_id * (order_num - 2), 0);
  Eigen::MatrixXd Pm; Pm.noalias() = basis * bz;
  std::memcpy(P, Pm.data(), 18 * sizeof(double));
}

// raw GJK witness vector between two vertex sets (rows x 3 col-major each)
void ref_gjk(const double* A, int na, const double* B, int nb, double* v) {
  std::vector<double*> pa(na), pb(nb);
  std::vector<double> ca(3 * na), cb(3 * nb);
  for (int i = 0; i < na; i++) { for (int j = 0; j < 3; j++) ca[3 * i + j] = A[j * na + i]; pa[i] = &ca[3 * i]; }
  for (int i = 0; i < nb; i++) { for (int j = 0; j < 3; j++) cb[3 * i + j] = B[j * nb + i]; pb[i] = &cb[3 * i]; }
  struct bd b1, b2; struct simplex s;
  b1.coord = pa.data(); b1.numpoints = na; b2.coord = pb.data(); b2.numpoints = nb; s.nvrtx = 0;
  double* c0 = gjk(b1, b2, &s);
  v[0] = c0[0]; v[1] = c0[1]; v[2] = c0[2];
}

int ref_kdop_dcd(const double* P, const double* q, double d) {
  Data Pm = map_mat(P, 6, 3); Data qm = map_mat(q, 1, 3);
  return CCD::KDOPDCD(Pm, qm, d);
}
int ref_self_kdop_dcd(const double* P0, const double* P1, double d) {
  return CCD::SelfKDOPDCD(map_mat(P0, 6, 3), map_mat(P1, 6, 3), d);
}
int ref_kdop_ccd(const double* P, const double* D, const double* q, double d, double t0, double t1) {
  return CCD::KDOPCCD(map_mat(P, 6, 3), map_mat(D, 6, 3), map_mat(q, 1, 3), d, t0, t1);
}
int ref_gjk_ccd(const double* P, const double* D, const double* q, double d, double t0, double t1) {
  return CCD::GJKCCD(map_mat(P, 6, 3), map_mat(D, 6, 3), map_mat(q, 1, 3), d, t0, t1);
}
int ref_self_kdop_ccd(const double* P0, const double* D0, const double* P1, const double* D1, double d,
                      double t0, double t1, double s0, double s1) {
  return CCD::SelfKDOPCCD(map_mat(P0, 6, 3), map_mat(D0, 6, 3), map_mat(P1, 6, 3), map_mat(D1, 6, 3), d, t0, t1, s0, s1);
}
int ref_self_gjk_ccd(const double* P0, const double* D0, const double* P1, const double* D1, double d,
                     double t0, double t1, double s0, double s1) {
  return CCD::SelfGJKCCD(map_mat(P0, 6, 3), map_mat(D0, 6, 3), map_mat(P1, 6, 3), map_mat(D1, 6, 3), d, t0, t1, s0, s1);
}
int ref_opengjk(const double* P, const double* q, double dist, double* c, double* d) {
  Eigen::Vector3d cv; double dv = 0;
  bool ok = Separate::opengjk(map_mat(P, 6, 3), map_mat(q, 1, 3), dist, cv, dv);
  c[0] = cv(0); c[1] = cv(1); c[2] = cv(2); *d = dv;
  return ok;
}
int ref_selfgjk(const double* P0, const double* P1, double dist, double* c, double* d) {
  Eigen::Vector3d cv; double dv = 0;
  bool ok = Separate::selfgjk(map_mat(P0, 6, 3), map_mat(P1, 6, 3), dist, cv, dv);
  c[0] = cv(0); c[1] = cv(1); c[2] = cv(2); *d = dv;
  return ok;
}
void ref_optimal_d(const double* P0, const double* P1, const double* c, double* d) {
  Eigen::Vector3d cv(c[0], c[1], c[2]);
  Optimal_plane::optimal_d(map_mat(P0, 6, 3), map_mat(P1, 6, 3), cv, *d);
}


// ---- persistent-plane mode ("optimal_plane": 1) ------------------------------------------------------------------------
// Optimal_plane::optimal_cd (Optimal_plane.h:160-293): Newton on the tangent angles of c, d tied to the point.
void ref_optimal_cd(const double* P, const double* q, double* c, double* d) {
  Eigen::Vector3d cv(c[0], c[1], c[2]);
  Optimal_plane::optimal_cd(map_mat(P, 6, 3), map_mat(q, 1, 3), cv, *d);
  c[0] = cv(0); c[1] = cv(1); c[2] = cv(2);
}
// Optimal_plane::self_optimal_cd (Optimal_plane.h:620-773): 3-variable (theta, phi, d) Newton, inter-robot plane.
void ref_self_optimal_cd(const double* P0, const double* P1, double* c, double* d) {
  Eigen::Vector3d cv(c[0], c[1], c[2]);
  Optimal_plane::self_optimal_cd(map_mat(P0, 6, 3), map_mat(P1, 6, 3), cv, *d);
  c[0] = cv(0); c[1] = cv(1); c[2] = cv(2);
}
// the persistent state of Main/admmPathPlanning3D.cpp:342-351 and Main/multiPathPlanning3D.cpp:450-464, emptied.
// dense N_tr x N_pts like the reference: only for small clouds.
void ref_reset_persistent_planes() {
#ifdef TRAJOPT_HOST_H   // built against the drop-in headers: the live planes are kept by the device context
  tob_host::Session& S = tob_host::Session::get();
  S.sync();
  S.check(tob_planes_reset(S.ctx()), "tob_planes_reset");
#endif
  int n_tr = subdivide_tree.size();
  size_t nv = g_vertex_list.size();
  is_seperate.assign(n_tr, std::vector<bool>(nv, false));
  seperate_c.assign(n_tr, std::vector<Eigen::Vector3d>(nv));
  seperate_d.assign(n_tr, std::vector<double>(nv));
  is_self_seperate.assign(n_tr, std::vector<std::vector<bool>>(uav_num, std::vector<bool>(uav_num, false)));
  self_seperate_c.assign(n_tr, std::vector<std::vector<Eigen::Vector3d>>(uav_num, std::vector<Eigen::Vector3d>(uav_num)));
  self_seperate_d.assign(n_tr, std::vector<std::vector<double>>(uav_num, std::vector<double>(uav_num)));
}
// live obstacle planes: (tr, point id) pairs with is_seperate set, in (tr, id) order; returns the count
long ref_live_planes(unsigned* tr, unsigned* id, double* c, double* d, long cap) {
#ifdef TRAJOPT_HOST_H
  {
    tob_host::Session& S = tob_host::Session::get();
    uint64_t total = 0;
    S.check(tob_live_planes(S.ctx(), tr, id, c, d, (uint64_t)cap, &total), "tob_live_planes");
    return (long)total;
  }
#endif
  long n = 0;
  for (size_t t = 0; t < is_seperate.size(); t++)
    for (size_t k = 0; k < is_seperate[t].size(); k++)
      if (is_seperate[t][k]) {
        if (n < cap) { tr[n] = t; id[n] = k; c[3 * n] = seperate_c[t][k](0); c[3 * n + 1] = seperate_c[t][k](1); c[3 * n + 2] = seperate_c[t][k](2); d[n] = seperate_d[t][k]; }
        n++;
      }
  return n;
}

// ---- planes -------------------------------------------------------------------
long ref_separate_plane(const double* spline, unsigned* off, double* c, double* d, long cap) {
  Quiet q;
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  Optimization3D_admm::separate_plane(map_mat(spline, trajectory_num, 3), g_vertex_list, cl, dl, *g_bvh);
  return planes_to_csr(cl, dl, off, c, d, cap);
}

// ---- energies -----------------------------------------------------------------
double ref_plane_barrier_energy(const double* spline, const unsigned* off, const double* c, const double* d) {
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  return Energy_admm::plane_barrier_energy(map_mat(spline, trajectory_num, 3), cl, dl);
}
double ref_bound_energy(const double* spline, double piece_time) {
  return Energy_admm::bound_energy(map_mat(spline, trajectory_num, 3), piece_time);
}
double ref_spline_energy(const double* spline, double piece_time, const double* p_slack, const double* t_slack,
                         const double* p_lambda, const double* t_lambda, const unsigned* off, const double* c,
                         const double* d) {
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  return Energy_admm::spline_energy(map_mat(spline, trajectory_num, 3), piece_time, map_mat(p_slack, 6 * piece_num, 3), ts,
                                    map_mat(p_lambda, 6 * piece_num, 3), tl, cl, dl);
}
double ref_slack_energy(const double* c_spline, double piece_time, const double* p_part, double t_part,
                        const double* p_lambda, double t_lambda) {
  return Energy_admm::slack_energy(map_mat(c_spline, 6, 3), piece_time, map_mat(p_part, 6, 3), t_part,
                                   map_mat(p_lambda, 6, 3), t_lambda);
}

// ---- gradients ----------------------------------------------------------------
// per piece blocks BEFORE the PSD projection: g0[19], h0[19x19 col-major]
void ref_local_spline_gradient(const double* spline, double piece_time, const double* p_slack, const double* t_slack,
                               const double* p_lambda, const double* t_lambda, const unsigned* off, const double* c,
                               const double* d, int sp_id, double* g0, double* h0) {
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Eigen::VectorXd g; Eigen::MatrixXd h;
  Gradient_admm::local_spline_gradient(map_mat(spline, trajectory_num, 3), piece_time, map_mat(p_slack, 6 * piece_num, 3), ts,
                                       map_mat(p_lambda, 6 * piece_num, 3), tl, cl, dl, g, h, sp_id);
  std::memcpy(g0, g.data(), 19 * sizeof(double));
  std::memcpy(h0, h.data(), 361 * sizeof(double));
}
// assembled: grad[3T+1], hessian[(3T+1)^2 col-major]
void ref_global_spline_gradient(const double* spline, double piece_time, const double* p_slack, const double* t_slack,
                                const double* p_lambda, const double* t_lambda, const unsigned* off, const double* c,
                                const double* d, double* grad, double* hess) {
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Eigen::VectorXd g; Eigen::MatrixXd h;
  Gradient_admm::global_spline_gradient(map_mat(spline, trajectory_num, 3), piece_time, map_mat(p_slack, 6 * piece_num, 3), ts,
                                        map_mat(p_lambda, 6 * piece_num, 3), tl, cl, dl, g, h);
  int n = 3 * trajectory_num + 1;
  std::memcpy(grad, g.data(), n * sizeof(double));
  std::memcpy(hess, h.data(), (size_t)n * n * sizeof(double));
}

// Optimization3D_admm::spline_descent_direction (Optimization3D_admm.h:400-503)
void ref_descent_direction(const double* spline, double piece_time, const double* p_slack, const double* t_slack,
                           const double* p_lambda, const double* t_lambda, const unsigned* off, const double* c,
                           const double* d, double* direction, double* t_direction, double* wolfe_out, double* gnorm_out) {
  Quiet q;
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Data dir; double td = 0;
  Optimization3D_admm::spline_descent_direction(map_mat(spline, trajectory_num, 3), dir, piece_time, td,
                                                map_mat(p_slack, 6 * piece_num, 3), ts, map_mat(p_lambda, 6 * piece_num, 3), tl,
                                                cl, dl);
  std::memcpy(direction, dir.data(), 3 * trajectory_num * sizeof(double));
  *t_direction = td; *wolfe_out = wolfe; *gnorm_out = gnorm;
}

// Optimization3D_multi::spline_descent_direction (dense LLT + eigen shift, Optimization3D_multi.h:659-752)
void ref_descent_direction_multi(const double* spline, double piece_time, const double* p_slack, const double* t_slack,
                                 const double* p_lambda, const double* t_lambda, const unsigned* off, const double* c,
                                 const double* d, double* direction, double* t_direction, double* wolfe_out,
                                 double* gnorm_add) {
  Quiet q;
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Data dir; double td = 0;
  double g_before = gnorm; gnorm = 0;
  Optimization3D_multi::spline_descent_direction(map_mat(spline, trajectory_num, 3), dir, piece_time, td,
                                                 map_mat(p_slack, 6 * piece_num, 3), ts, map_mat(p_lambda, 6 * piece_num, 3), tl,
                                                 cl, dl);
  *gnorm_add = gnorm; gnorm = g_before;
  std::memcpy(direction, dir.data(), 3 * trajectory_num * sizeof(double));
  *t_direction = td; *wolfe_out = wolfe;
}

// ---- CCD step -------------------------------------------------------------------
double ref_position_step(const double* spline, const double* direction) {
  Quiet q;
  return Step::position_step(map_mat(spline, trajectory_num, 3), map_mat(direction, trajectory_num, 3), g_vertex_list, *g_bvh);
}

void ref_self_step(const double* splines, const double* directions, int u, double* steps) {
  Quiet q;
  std::vector<Data> sl, dl;
  for (int i = 0; i < u; i++) {
    sl.push_back(map_mat(splines + (size_t)3 * trajectory_num * i, trajectory_num, 3));
    dl.push_back(map_mat(directions + (size_t)3 * trajectory_num * i, trajectory_num, 3));
  }
  std::vector<double> st;
  BVH b;
  Step::self_step(sl, dl, st, b);
  for (int i = 0; i < u; i++) steps[i] = st[i];
}

double ref_couple_self_step(const double* splines, const double* directions, int u) {
  Quiet q;
  std::vector<Data> sl, dl;
  for (int i = 0; i < u; i++) {
    sl.push_back(map_mat(splines + (size_t)3 * trajectory_num * i, trajectory_num, 3));
    dl.push_back(map_mat(directions + (size_t)3 * trajectory_num * i, trajectory_num, 3));
  }
  double step = 1.0;
  BVH b;
  Step::couple_self_step(sl, dl, step, b);
  return step;
}

// ---- slack / dual update ---------------------------------------------------------
void ref_update_slack_lambda(const double* spline, double piece_time, double* p_slack, double* t_slack, double* p_lambda,
                             double* t_lambda) {
  Quiet q;
  Data ps = map_mat(p_slack, 6 * piece_num, 3), pl = map_mat(p_lambda, 6 * piece_num, 3);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Optimization3D_admm::update_slack_lambda(map_mat(spline, trajectory_num, 3), piece_time, ps, ts, pl, tl);
  std::memcpy(p_slack, ps.data(), 18 * piece_num * sizeof(double));
  std::memcpy(p_lambda, pl.data(), 18 * piece_num * sizeof(double));
  std::memcpy(t_slack, ts.data(), piece_num * sizeof(double));
  std::memcpy(t_lambda, tl.data(), piece_num * sizeof(double));
}

// ---- whole iterations --------------------------------------------------------------
// Optimization3D_admm::optimization (Optimization3D_admm.h:29-67); state in/out.
void ref_optimization(double* spline, double* piece_time, double* p_slack, double* t_slack, double* p_lambda,
                      double* t_lambda, double* gnorm_out) {
  Quiet q;
  Data sp = map_mat(spline, trajectory_num, 3);
  Data ps = map_mat(p_slack, 6 * piece_num, 3), pl = map_mat(p_lambda, 6 * piece_num, 3);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Optimization3D_admm::optimization(sp, *piece_time, ps, ts, pl, tl, g_vertex_list, *g_bvh);
  std::memcpy(spline, sp.data(), 3 * trajectory_num * sizeof(double));
  std::memcpy(p_slack, ps.data(), 18 * piece_num * sizeof(double));
  std::memcpy(p_lambda, pl.data(), 18 * piece_num * sizeof(double));
  std::memcpy(t_slack, ts.data(), piece_num * sizeof(double));
  std::memcpy(t_lambda, tl.data(), piece_num * sizeof(double));
  *gnorm_out = gnorm;
  iter++;
}

// multi: arrays are u blocks back to back. coupled=0 -> optimization_decouple (piece_time[u]),
// coupled=1 -> optimization (piece_time[0] shared).
void ref_optimization_multi(int coupled, int u, double* splines, double* piece_time, double* p_slack, double* t_slack,
                            double* p_lambda, double* t_lambda, double* gnorm_out) {
  Quiet q;
  std::vector<Data> sl(u), psl(u), pll(u);
  std::vector<Eigen::VectorXd> tsl(u), tll(u);
  std::vector<double> ptl(u);
  size_t ns = 3 * trajectory_num, np = 18 * piece_num;
  for (int i = 0; i < u; i++) {
    sl[i] = map_mat(splines + ns * i, trajectory_num, 3);
    psl[i] = map_mat(p_slack + np * i, 6 * piece_num, 3);
    pll[i] = map_mat(p_lambda + np * i, 6 * piece_num, 3);
    tsl[i] = Eigen::Map<const Eigen::VectorXd>(t_slack + (size_t)piece_num * i, piece_num);
    tll[i] = Eigen::Map<const Eigen::VectorXd>(t_lambda + (size_t)piece_num * i, piece_num);
    ptl[i] = piece_time[coupled ? 0 : i];
  }
  if (coupled) {
    double pt = piece_time[0];
    Optimization3D_multi::optimization(sl, pt, psl, tsl, pll, tll, g_vertex_list, *g_bvh);
    piece_time[0] = pt;
  } else {
    Optimization3D_multi::optimization_decouple(sl, ptl, psl, tsl, pll, tll, g_vertex_list, *g_bvh);
    for (int i = 0; i < u; i++) piece_time[i] = ptl[i];
  }
  for (int i = 0; i < u; i++) {
    std::memcpy(splines + ns * i, sl[i].data(), ns * sizeof(double));
    std::memcpy(p_slack + np * i, psl[i].data(), np * sizeof(double));
    std::memcpy(p_lambda + np * i, pll[i].data(), np * sizeof(double));
    std::memcpy(t_slack + (size_t)piece_num * i, tsl[i].data(), piece_num * sizeof(double));
    std::memcpy(t_lambda + (size_t)piece_num * i, tll[i].data(), piece_num * sizeof(double));
  }
  *gnorm_out = gnorm;
  iter++;
}

// inter-robot planes only: Optimization3D_multi::separate_self appended onto EMPTY lists.
// out: per (robot, tr) CSR over u*n_tr rows.
long ref_separate_self(const double* splines, int u, unsigned* off, double* c, double* d, long cap) {
  Quiet q;
  int n_tr = subdivide_tree.size();
  std::vector<Data> sl(u);
  for (int i = 0; i < u; i++) sl[i] = map_mat(splines + (size_t)3 * trajectory_num * i, trajectory_num, 3);
  std::vector<std::vector<std::vector<Eigen::Vector3d>>> cl(u);
  std::vector<std::vector<std::vector<double>>> dl(u);
  for (int i = 0; i < u; i++) { cl[i].resize(n_tr); dl[i].resize(n_tr); }
  BVH b;
  Optimization3D_multi::separate_self(sl, cl, dl, b);
  long n = 0; off[0] = 0;
  for (int i = 0; i < u; i++)
    for (int t = 0; t < n_tr; t++) {
      for (size_t k = 0; k < cl[i][t].size(); k++) {
        if (n < cap) { c[3 * n] = cl[i][t][k](0); c[3 * n + 1] = cl[i][t][k](1); c[3 * n + 2] = cl[i][t][k](2); d[n] = dl[i][t][k]; }
        n++;
      }
      off[i * n_tr + t + 1] = (unsigned)n;
    }
  return n;
}

// ---- function-level entry points behind the remaining C-ABI rows (tests/test_gpu_abi_entries.py) ---------------------
// Gradient_admm::local_plane_barrier_gradient (Gradient_admm.h:331-407): one sub-segment against its plane list
void ref_local_plane_barrier_gradient(const double* spline, int tr_id, const double* c, const double* d, int n, double* grad18,
                                      double* hess324) {
  std::vector<Eigen::Vector3d> cl; std::vector<double> dl;
  for (int k = 0; k < n; k++) { cl.push_back(Eigen::Vector3d(c[3 * k], c[3 * k + 1], c[3 * k + 2])); dl.push_back(d[k]); }
  Eigen::VectorXd g; Eigen::MatrixXd h;
  Gradient_admm::local_plane_barrier_gradient(tr_id, map_mat(spline, trajectory_num, 3), cl, dl, g, h);
  std::memcpy(grad18, g.data(), 18 * sizeof(double));
  std::memcpy(hess324, h.data(), 324 * sizeof(double));
}
// Gradient_admm::local_bound_gradient (Gradient_admm.h:409-572)
void ref_local_bound_gradient(const double* spline, int tr_id, double piece_time, double* grad18, double* hess324, double* g_t,
                              double* h_t, double* partgrad18) {
  Eigen::VectorXd g, pg; Eigen::MatrixXd h;
  Gradient_admm::local_bound_gradient(tr_id, map_mat(spline, trajectory_num, 3), piece_time, g, h, *g_t, *h_t, pg);
  std::memcpy(grad18, g.data(), 18 * sizeof(double));
  std::memcpy(hess324, h.data(), 324 * sizeof(double));
  std::memcpy(partgrad18, pg.data(), 18 * sizeof(double));
}
// Gradient_admm::slack_gradient (Gradient_admm.h:574-622): grad[19], hessian[19x19 col-major]
void ref_slack_gradient(const double* c_spline, double piece_time, const double* p_part, double t_part, const double* p_lambda,
                        double t_lambda, double* grad19, double* hess361) {
  Eigen::VectorXd g; Eigen::MatrixXd h;
  Gradient_admm::slack_gradient(map_mat(c_spline, 6, 3), piece_time, map_mat(p_part, 6, 3), t_part, map_mat(p_lambda, 6, 3),
                                t_lambda, g, h);
  std::memcpy(grad19, g.data(), 19 * sizeof(double));
  std::memcpy(hess361, h.data(), 361 * sizeof(double));
}
// Energy_admm::dynamic_energy (Energy_admm.h:199-215), Gradient_admm::dynamic_gradient (Gradient_admm.h:633-671)
double ref_dynamic_energy(const double* p_part, double t_part) { return Energy_admm::dynamic_energy(map_mat(p_part, 6, 3), t_part); }
void ref_dynamic_gradient(const double* p_part, double t_part, double* grad18, double* hess324, double* g_t, double* h_t,
                          double* partgrad18) {
  Eigen::VectorXd g, pg; Eigen::MatrixXd h;
  Gradient_admm::dynamic_gradient(map_mat(p_part, 6, 3), t_part, g, h, *g_t, *h_t, pg);
  std::memcpy(grad18, g.data(), 18 * sizeof(double));
  std::memcpy(hess324, h.data(), 324 * sizeof(double));
  std::memcpy(partgrad18, pg.data(), 18 * sizeof(double));
}
// Optimization3D_admm::spline_line_search (Optimization3D_admm.h:505-557): CCD bound from Step::position_step, then Armijo
// against the given planes with the global `wolfe`; spline / piece_time in-out
void ref_line_search(double* spline, double* piece_time, const double* direction, double t_direction, double wolfe_in,
                     const double* p_slack, const double* t_slack, const double* p_lambda, const double* t_lambda,
                     const unsigned* off, const double* c, const double* d) {
  Quiet q;
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Data sp = map_mat(spline, trajectory_num, 3);
  double pt = *piece_time;
  wolfe = wolfe_in;
  Optimization3D_admm::spline_line_search(sp, map_mat(direction, trajectory_num, 3), pt, t_direction,
                                          map_mat(p_slack, 6 * piece_num, 3), ts, map_mat(p_lambda, 6 * piece_num, 3), tl,
                                          g_vertex_list, *g_bvh, cl, dl);
  std::memcpy(spline, sp.data(), 3 * trajectory_num * sizeof(double));
  *piece_time = pt;
}
// Optimization3D_multi::spline_line_search (Optimization3D_multi.h:754-811): the caller's step bound, in-out
void ref_line_search_multi(double* spline, double* piece_time, const double* direction, double t_direction, double wolfe_in,
                           const double* p_slack, const double* t_slack, const double* p_lambda, const double* t_lambda,
                           const unsigned* off, const double* c, const double* d, double* step_io) {
  Quiet q;
  std::vector<std::vector<Eigen::Vector3d>> cl; std::vector<std::vector<double>> dl;
  planes_from_csr(off, c, d, subdivide_tree.size(), cl, dl);
  Eigen::VectorXd ts = Eigen::Map<const Eigen::VectorXd>(t_slack, piece_num);
  Eigen::VectorXd tl = Eigen::Map<const Eigen::VectorXd>(t_lambda, piece_num);
  Data sp = map_mat(spline, trajectory_num, 3);
  double pt = *piece_time, st = *step_io;
  wolfe = wolfe_in;
  Optimization3D_multi::spline_line_search(sp, map_mat(direction, trajectory_num, 3), pt, t_direction,
                                           map_mat(p_slack, 6 * piece_num, 3), ts, map_mat(p_lambda, 6 * piece_num, 3), tl, cl, dl, st);
  std::memcpy(spline, sp.data(), 3 * trajectory_num * sizeof(double));
  *piece_time = pt; *step_io = st;
}
// BVH::EdgeCollision (BVH.cpp:95-133): ids of all points within d of the box of a 2-point edge (2x3 col-major)
long ref_edge_collision(const double* edge, double d, unsigned* ids, long cap) {
  std::vector<unsigned int> cp;
  g_bvh->EdgeCollision(map_mat(edge, 2, 3), cp, d);
  for (size_t k = 0; k < cp.size() && (long)k < cap; k++) ids[k] = cp[k];
  return (long)cp.size();
}
// CCD::GJKDCD (CCD.h:17-114)
int ref_gjk_dcd(const double* A, int na, const double* B, int nb, double d) {
  return CCD::GJKDCD(map_mat(A, na, 3), map_mat(B, nb, 3), d) ? 1 : 0;
}

// Mesh::readOBJ (HighOrderCCD/Utils/CCDUtils.h:317-391): returns the vertex count, fills V (n x 3 col-major) up to cap rows
long ref_read_obj(const char* path, double* V, long cap) {
  Eigen::MatrixXd M;
  Mesh::readOBJ(std::string(path), M);
  const long n = M.rows();
  for (long i = 0; i < n && i < cap; i++) for (int k = 0; k < 3; k++) V[i + cap * k] = M(i, k);
  return n;
}

}  // extern "C"
