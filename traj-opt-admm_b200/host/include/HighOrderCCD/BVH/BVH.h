// Drop-in for HighOrderCCD/BVH/BVH.h (reference :11-41) + BVH.cpp: same class name, method names and argument meaning;
// the incremental aabb::Tree (BVH/src/AABB.{h,cc}) is replaced by the GPU LBVH behind the C ABI, so the public tree
// members tr_tree / ob_tree / pc_tree of the reference do not exist here (nothing outside BVH.cpp reads them).
// Candidate lists come back sorted by id; the reference returns its tree's DFS order (the contract is the set).
#ifndef BVH_H
#define BVH_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/CCD/CCD.h"
#include "trajopt_host.h"

PRJ_BEGIN

class BVH {
 public:
  typedef std::vector<std::tuple<int, std::pair<double, double>, Eigen::MatrixXd>> SubdivideTree;
  typedef Eigen::MatrixXd Data;
  typedef std::pair<unsigned int, unsigned int> id_pair;

  BVH() {}
  ~BVH() {}

  // reference BVH.cpp:15-51 is dead code (triangle-mesh obstacles are never built by Main)
  void InitObstacle(const Eigen::MatrixXd&, const Eigen::MatrixXi&) {
    throw std::runtime_error("BVH::InitObstacle: triangle-mesh obstacles are not part of the B200 hot path");
  }

  // BVH.cpp:53-92 -> Morton-sorted LBVH on the device
  void InitPointcloud(const Eigen::MatrixXd& V) {
    tob_host::Session& S = tob_host::Session::get();
    if (V.cols() != 3) throw std::runtime_error("BVH::InitPointcloud: V must be n x 3");
    Eigen::MatrixXd Vc = V;   // contiguous column-major copy (V may be an expression)
    S.upload_cloud(Vc.data(), (uint32_t)Vc.rows());     // replicated on every GPU of the session
  }

  // BVH.cpp:95-133: points within d of the box of a 2-point edge
  void EdgeCollision(const Data& edge, std::vector<unsigned int>& collision_pair, double d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
      lo[k] = INFINITY; hi[k] = -INFINITY;
      for (int j = 0; j < 2; j++) { double v = edge(j, k); if (v < lo[k]) lo[k] = v; if (v > hi[k]) hi[k] = v; }
    }
    uint64_t total = 0;
    std::vector<uint32_t> ids(4096);
    S.check(tob_box_query(S.ctx(), lo, hi, d, ids.data(), ids.size(), &total), "tob_box_query");
    if (total > ids.size()) {
      ids.resize(total);
      S.check(tob_box_query(S.ctx(), lo, hi, d, ids.data(), ids.size(), &total), "tob_box_query");
    }
    collision_pair.assign(ids.begin(), ids.begin() + total);
  }

  // BVH.cpp:136-147 is dead code
  void SelfEdgeCollision(const std::vector<Data>&, std::vector<id_pair>&, double) {
    throw std::runtime_error("BVH::SelfEdgeCollision is not part of the B200 hot path");
  }

  // BVH.cpp:149-192
  void DCDCollision(const Data& spline, std::vector<std::vector<unsigned int>>& collision_pairs, double d) {
    query(spline, nullptr, collision_pairs, d);
  }
  // BVH.cpp:195-249
  void CCDCollision(const Data& spline, const Data& direction, std::vector<std::vector<unsigned int>>& collision_pairs, double d) {
    query(spline, &direction, collision_pairs, d);
  }
  // BVH.cpp:252-286
  void SelfDCDCollision(const std::vector<Data>& P, std::vector<id_pair>& collision_pair, double d) { self_query(P, nullptr, collision_pair, d); }
  // BVH.cpp:289-329
  void SelfCCDCollision(const std::vector<Data>& P, const std::vector<Data>& D, std::vector<id_pair>& collision_pair, double d) {
    self_query(P, &D, collision_pair, d);
  }

 private:
  void query(const Data& spline, const Data* direction, std::vector<std::vector<unsigned int>>& out, double d) {
    tob_host::Session& S = tob_host::Session::get();
    S.sync(); S.ensure_cloud();
    const int n_tr = piece_num * res;
    std::vector<uint32_t> off(n_tr + 1), ids(1 << 16);
    uint64_t total = 0;
    for (int pass = 0; pass < 2; pass++) {
      int rc = direction ? tob_broadphase_ccd(S.ctx(), spline.data(), direction->data(), 1, d, off.data(), ids.data(), ids.size(), &total)
                         : tob_broadphase_dcd(S.ctx(), spline.data(), 1, d, off.data(), ids.data(), ids.size(), &total);
      S.check(rc, "tob_broadphase");
      if (total <= ids.size()) break;
      ids.resize(total);
    }
    out.assign(n_tr, std::vector<unsigned int>());
    for (int r = 0; r < n_tr; r++) out[r].assign(ids.begin() + off[r], ids.begin() + off[r + 1]);
  }
  void self_query(const std::vector<Data>& P, const std::vector<Data>* D, std::vector<id_pair>& out, double d) {
    tob_host::Session& S = tob_host::Session::get();
    const int u = (int)P.size();
    std::vector<double> p((size_t)18 * u), dd(D ? (size_t)18 * u : 0);
    for (int i = 0; i < u; i++) {
      if (P[i].rows() != 6 || P[i].cols() != 3) throw std::runtime_error("BVH::Self*Collision: control polygons must be 6x3");
      std::memcpy(&p[(size_t)18 * i], P[i].data(), 18 * sizeof(double));
      if (D) std::memcpy(&dd[(size_t)18 * i], (*D)[i].data(), 18 * sizeof(double));
    }
    std::vector<uint32_t> pairs((size_t)u * (u > 1 ? u - 1 : 1));   // 2 * u(u-1)/2
    uint64_t total = 0;
    S.check(tob_self_broadphase(S.ctx(), p.data(), D ? dd.data() : nullptr, u, d, pairs.data(), pairs.size() / 2, &total), "tob_self_broadphase");
    out.clear();
    for (uint64_t i = 0; i < total; i++) out.push_back(id_pair(pairs[2 * i], pairs[2 * i + 1]));
  }
};

PRJ_END

#endif
