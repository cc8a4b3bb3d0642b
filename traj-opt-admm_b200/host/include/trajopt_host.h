// trajopt_host.h -- glue between the reference-shaped C++ entry points (the shadow headers next to this file) and the
// C ABI of the CUDA library (include/trajopt_b200.h).
//
// The reference keeps every solver parameter in namespace-scope globals (HighOrderCCD/Utils/CCDUtils.h:36-82, defined in
// CCDUtils.cpp:5-44) that Main/*.cpp fill after start-up; that file pair is host set-up code and is used UNCHANGED from the
// reference tree.  Session::sync() mirrors the globals the hot path reads into the device context whenever they change:
//   piece_num, res, uav_num, is_optimal_plane, lambda, margin, offset, mu, vel_limit, acc_limit, ks, kt   -> tob_set_params
//   subdivide_tree (basis + parameter range), convert_list, M_dynamic, kdop_matrix                      -> tob_set_tables
// One process-wide session (Main constructs only a BVH object and passes it around).  It owns one
// context per GPU: TRAJOPT_B200_GPUS=N (default 1) makes the multi-UAV iterations of Main/multiPathPlanning3D.cpp run with
// the robots sharded over N GPUs -- one context and one host thread per GPU, the per-iteration exchange on NCCL inside the
// library (tob_nccl_init_all) -- without any change to the caller; the result is bitwise the one of a single GPU.  Function-
// level entry points always use the first context.
// Errors of the C ABI become std::runtime_error: there is no CPU fallback.
#ifndef TRAJOPT_HOST_H
#define TRAJOPT_HOST_H

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "trajopt_b200.h"

namespace tob_host {

typedef Eigen::MatrixXd Data;
typedef std::vector<std::vector<Eigen::Vector3d>> CLists;
typedef std::vector<std::vector<double>> DLists;

class Session {
 public:
  static Session& get() {
    static Session s;
    return s;
  }
  tob_ctx* ctx() { return ctx_; }
  int gpus() const { return (int)all_.size(); }

  void check(int rc, const char* what) {
    if (rc) throw std::runtime_error(std::string(what) + ": " + tob_last_error(ctx_));
  }

  // BVH::InitPointcloud: the cloud (and its LBVH) is replicated on every GPU of the session
  void upload_cloud(const double* V, uint32_t n) {
    for (size_t g = 0; g < all_.size(); g++)
      if (tob_cloud_upload(all_[g], V, n)) throw std::runtime_error(std::string("tob_cloud_upload: ") + tob_last_error(all_[g]));
  }

  // one multi-UAV ADMM iteration over all GPUs of the session: every context gets all the states (it uses the robots it
  // owns) and writes back the robots it owns; the threads only exist for the duration of the call
  double optimization(tob_state* st, int n_robots, int mode) {
    const int G = (int)all_.size();
    if (G == 1 || n_robots < G) {
      double gn = 0;
      check(tob_optimization(ctx_, st, n_robots, mode, &gn), "tob_optimization");
      return gn;
    }
    if (!sharded_) {
      if (tob_nccl_init_all(all_.data(), G)) throw std::runtime_error(std::string("tob_nccl_init_all: ") + tob_last_error(all_[0]));
      sharded_ = true;
    }
    std::vector<int> rc(G, 0);
    std::vector<double> gn(G, 0.0);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
      th.emplace_back([&, g]() { rc[g] = tob_optimization(all_[g], st, n_robots, mode, &gn[g]); });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
      if (rc[g]) throw std::runtime_error(std::string("tob_optimization (GPU ") + std::to_string(g) + "): " + tob_last_error(all_[g]));
    return gn[0];
  }

  // globals -> device (cheap compare; uploads only on change)
  void sync() {
    using namespace HighOrderCCD;
    tob_params p;
    std::memset(&p, 0, sizeof(p));
    p.piece_num = piece_num; p.res = res; p.uav_num = uav_num > 0 ? uav_num : 1; p.optimal_plane = is_optimal_plane ? 1 : 0;
    p.lambda = lambda; p.margin = margin; p.offset = offset; p.mu = mu;
    p.vel_limit = vel_limit; p.acc_limit = acc_limit; p.ks = ks; p.kt = kt;
    if (!have_params_ || std::memcmp(&p, &last_, sizeof(p)) != 0) {
      for (size_t g = 0; g < all_.size(); g++)
        if (tob_set_params(all_[g], &p)) throw std::runtime_error(std::string("tob_set_params: ") + tob_last_error(all_[g]));
      last_ = p; have_params_ = true; tables_.clear();
    }
    const size_t n_tr = (size_t)piece_num * res;
    if (subdivide_tree.size() != n_tr || (int)convert_list.size() != piece_num || M_dynamic.size() != 36 ||
        kdop_matrix.rows() != 3 || kdop_matrix.cols() != 49)
      throw std::runtime_error("trajopt_host: subdivide_tree / convert_list / M_dynamic / kdop_matrix are not initialised "
                               "(Main/admmPathPlanning3D.cpp:294-338,403-410)");
    std::vector<double> t(n_tr * 37 + (size_t)piece_num * 36 + 36 + 147);
    double* basis = t.data(); double* weight = basis + n_tr * 36; double* conv = weight + n_tr; double* md = conv + (size_t)piece_num * 36;
    double* kd = md + 36;
    for (size_t r = 0; r < n_tr; r++) {
      const Eigen::MatrixXd& b = std::get<2>(subdivide_tree[r]);
      if (b.rows() != 6 || b.cols() != 6) throw std::runtime_error("trajopt_host: subdivide_tree basis must be 6x6 (order_num=5)");
      std::memcpy(basis + 36 * r, b.data(), 36 * sizeof(double));
      weight[r] = std::get<1>(subdivide_tree[r]).second - std::get<1>(subdivide_tree[r]).first;
    }
    for (int i = 0; i < piece_num; i++) std::memcpy(conv + 36 * i, convert_list[i].data(), 36 * sizeof(double));
    std::memcpy(md, M_dynamic.data(), 36 * sizeof(double));
    std::memcpy(kd, kdop_matrix.data(), 147 * sizeof(double));
    if (t != tables_) {
      for (size_t g = 0; g < all_.size(); g++)
        if (tob_set_tables(all_[g], basis, weight, conv, md, kd)) throw std::runtime_error(std::string("tob_set_tables: ") + tob_last_error(all_[g]));
      tables_.swap(t);
    }
  }

  // the reference queries an empty tree when BVH::InitPointcloud was never called (init_ob = 0): no candidates.
  // The device path wants a cloud, so give it one point that can never be a candidate.
  void ensure_cloud() {
    if (tob_cloud_size(ctx_) == 0) {
      const double far_away[3] = {1e300, 1e300, 1e300};
      upload_cloud(far_away, 1);
    }
  }

 private:
  Session() {
    const char* dev = std::getenv("TRAJOPT_B200_DEVICE");
    const char* ng = std::getenv("TRAJOPT_B200_GPUS");
    const int first = dev ? std::atoi(dev) : 0, n = ng && std::atoi(ng) > 1 ? std::atoi(ng) : 1;
    for (int g = 0; g < n; g++) {
      tob_ctx* c = nullptr;
      if (tob_ctx_create(first + g, &c)) throw std::runtime_error(std::string("tob_ctx_create: ") + tob_last_error(nullptr));
      all_.push_back(c);
    }
    ctx_ = all_[0];
  }
  ~Session() { for (size_t g = 0; g < all_.size(); g++) tob_ctx_destroy(all_[g]); }
  Session(const Session&);
  tob_ctx* ctx_ = nullptr;            // first context: function-level entry points
  std::vector<tob_ctx*> all_;         // one context per GPU
  bool sharded_ = false;
  tob_params last_;
  bool have_params_ = false;
  std::vector<double> tables_;
};

// ragged plane lists of ONE robot -> CSR (c row-major xyz per plane)
struct PlaneCSR {
  std::vector<uint32_t> off;
  std::vector<double> c, d;
};
inline void planes_to_csr(const CLists& cl, const DLists& dl, PlaneCSR& out) {
  const size_t rows = cl.size();
  out.off.assign(rows + 1, 0u);
  out.c.clear(); out.d.clear();
  for (size_t r = 0; r < rows; r++) {
    for (size_t k = 0; k < cl[r].size(); k++) {
      out.c.push_back(cl[r][k](0)); out.c.push_back(cl[r][k](1)); out.c.push_back(cl[r][k](2));
      out.d.push_back(dl[r][k]);
    }
    out.off[r + 1] = (uint32_t)out.d.size();
  }
  if (out.c.empty()) { out.c.resize(3); out.d.resize(1); }   // never hand the ABI a null pointer
}
inline void csr_to_planes(const uint32_t* off, const double* c, const double* d, size_t rows, CLists& cl, DLists& dl, bool append) {
  if (!append) { cl.assign(rows, std::vector<Eigen::Vector3d>()); dl.assign(rows, std::vector<double>()); }
  for (size_t r = 0; r < rows; r++)
    for (uint32_t k = off[r]; k < off[r + 1]; k++) {
      cl[r].push_back(Eigen::Vector3d(c[3 * k], c[3 * k + 1], c[3 * k + 2]));
      dl[r].push_back(d[k]);
    }
}

// install the caller's plane lists of robot 0 as the resident set
inline void set_planes(const CLists& cl, const DLists& dl) {
  Session& S = Session::get();
  if ((int)cl.size() != HighOrderCCD::piece_num * HighOrderCCD::res) throw std::runtime_error("trajopt_host: c_lists must have piece_num*res rows");
  PlaneCSR p;
  planes_to_csr(cl, dl, p);
  S.check(tob_set_planes(S.ctx(), 1, p.off.data(), p.c.data(), p.d.data()), "tob_set_planes");
}

// tob_state over caller-owned Eigen storage (column-major, contiguous)
struct StateView {
  double pt;
  tob_state st;
  StateView(const Data& spline, double piece_time, const Data& p_slack, const Eigen::VectorXd& t_slack, const Data& p_lambda,
            const Eigen::VectorXd& t_lambda)
      : pt(piece_time) {
    st.spline = const_cast<double*>(spline.data()); st.piece_time = &pt;
    st.p_slack = const_cast<double*>(p_slack.data()); st.t_slack = const_cast<double*>(t_slack.data());
    st.p_lambda = const_cast<double*>(p_lambda.data()); st.t_lambda = const_cast<double*>(t_lambda.data());
  }
};

inline void check_state_shapes(const Data& spline, const Data& p_slack, const Eigen::VectorXd& t_slack, const Data& p_lambda,
                               const Eigen::VectorXd& t_lambda) {
  using namespace HighOrderCCD;
  const int T = 6 + 3 * (piece_num - 1);
  if (spline.rows() != T || spline.cols() != 3 || p_slack.rows() != 6 * piece_num || p_slack.cols() != 3 ||
      p_lambda.rows() != 6 * piece_num || p_lambda.cols() != 3 || t_slack.size() != piece_num || t_lambda.size() != piece_num)
    throw std::runtime_error("trajopt_host: state arrays do not match piece_num");
}

}  // namespace tob_host
#endif
