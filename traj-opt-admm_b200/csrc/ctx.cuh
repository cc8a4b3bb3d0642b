// ctx.cuh -- context object behind the C ABI (include/trajopt_b200.h) and shared launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/trajopt_b200.h"

#define TOB_MAX_LEVELS 8
#define TOB_LS_TRIALS 9      // line-search trial points per launch: index 0 = current point, 1..8 = ladder rungs
#define TOB_LADDER 400       // longest 0.8^k ladder the CCD kernels will walk
#define TOB_KF_ROW 100       // floats per row of the single-precision 49-DOP filter: 49 x (below, above) thresholds, allowance, pad

struct tob_ctx;

namespace tob {

// bumped whenever a device buffer is (re)allocated: a captured CUDA graph holds raw pointers and must be rebuilt
inline std::atomic<unsigned long long>& alloc_generation() {   // process-wide: contexts of several host threads share it
  static std::atomic<unsigned long long> g{0};
  return g;
}

// TRAJOPT_B200_POISON=1 (set by tests/conftest.py): every new device allocation is filled with 0xFF bytes (NaN doubles,
// 0xffffffff indices), so that a read of memory nobody wrote shows up in the tests instead of depending on what the
// allocator hands out (a fresh process gets zero pages, a long-lived one does not)
inline bool poison_allocations() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("TRAJOPT_B200_POISON"); on = (e && e[0] && e[0] != '0') ? 1 : 0; }
  return on == 1;
}

// growable device buffer (contents are NOT preserved on growth); owns its allocation (move-only)
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  cudaError_t ensure(size_t n) {
    if (n <= cap && p) return cudaSuccess;
    size_t want = n + n / 4 + 256;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    alloc_generation()++;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    if (e == cudaSuccess && poison_allocations()) e = cudaMemset(p, 0xff, want * sizeof(T));
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Device-resident counts and flags.  The iteration never reads a size back to the host in the middle: kernels take
// their trip counts from here, buffers have a fixed capacity, and an overflow / an unfinished line search turns the
// state-changing kernels at the end of the iteration into no-ops; the host looks at this struct ONCE per iteration.
#define TOB_OVF_CAND 1u      // broadphase produced more candidates than cand_cap
#define TOB_OVF_LIVE 8u      // persistent-plane mode: live planes + new planes exceed live_cap
#define TOB_OVF_REMOTE 16u   // sharded: another rank overflowed (every rank repeats the iteration, only the flagged ones grow)
#define TOB_OVF_RETRY (TOB_OVF_CAND | TOB_OVF_LIVE | TOB_OVF_REMOTE)   // the iteration is repeated: whatever ran on the incomplete plane set is void
#define TOB_ERR_SOLVE 32u    // Newton system not positive definite (Cholesky pivot / Schur complement <= 0): nothing is committed
#define TOB_LS_MAXROUNDS 16  // most Armijo rounds launched ahead of the host (the count is a run-time choice, see ls_policy)
struct DevCounts {
  uint32_t n_cand;           // candidates of the last broadphase fill (may exceed the capacity -> overflow)
  uint32_t n_planes;         // planes of the last pack
  uint32_t n_planes_ob;      // ... of which obstacle planes
  uint32_t bp_items;         // item records handed out by the broadphase count pass (reset by the fill pass)
  uint32_t n_en_items;       // listed virtual warps (v >= 1) of the rows with more than 256 planes (barrier.cu: k_en_items)
  uint32_t overflow;         // TOB_OVF_* bits, sticky until the host clears them
  int32_t ls_pending[TOB_LS_MAXROUNDS + 1];   // robots still backtracking after Armijo round r (last entry: host scratch)
  int32_t ls_rounds;                          // rounds launched ahead in this iteration (written by k_ls_init)
  uint32_t iters_done;       // iterations fully committed (apply step + slack update ran)
  uint32_t np_next;          // next chunk of candidates a CTA of k_narrow takes (reset by k_np_top, which runs behind it)
  unsigned long long dcd_candidates, planes, ccd_candidates, energy_plane_evals, barrier_terms;   // cumulative since reset
  // persistent-plane mode ("optimal_plane": 1)
  uint32_t n_live;           // live (row, point) obstacle planes
  uint32_t n_live_next;      // n_live + new planes of this pass (written by the merge, consumed by the refinement)
  uint32_t n_new;            // planes accepted for pairs that were not live yet
  uint32_t opt_capped;       // plane refinements that hit a loop cap (the reference's loops are unbounded)
  // counted work of the narrowphase / CCD kernels (cumulative since reset; bench.py turns them into algorithmic flops)
  unsigned long long np_kdop_groups;   // 7-axis groups of the 49-DOP gate really evaluated by k_narrow
  unsigned long long np_gjk_iters;     // GJK(6,1) rounds (support + sub-algorithm) run by k_narrow
  unsigned long long ccd_gjk_iters;    // GJK(12,1) rounds run by the CCD ladder of k_bp_ccd
  unsigned long long ccd_kdop_pass;    // swept candidates that passed the swept 49-DOP gate (each runs >= 1 ladder rung)
  unsigned long long np_kdop_exact;    // axes of the 49-DOP gate the single-precision filter could not decide (re-tested in FP64)
  unsigned long long np_band;          // pairs whose GJK distance fell inside the band where the rest of the gate had to be evaluated
  unsigned long long ls_hist[8];       // decoupled line searches by the ladder rung that was accepted: 1, 2, .. 7, 8 and deeper
  unsigned long long ls_skipped;       // ladder rungs not evaluated: a velocity / acceleration bound is violated for certain there (k_ls_bound_mask)
};

// per-row (robot x sub-segment) geometry produced by segments.cu, indexed by GLOBAL row = robot*n_tr + tr
struct RowGeom {
  DBuf<double> P;      // rows x 18  (6x3 col-major: P[j + 6*axis])
  DBuf<double> D;      // rows x 18  direction control points (CCD)
  DBuf<double> box;    // rows x 6   lo xyz, hi xyz
  DBuf<double> klo;    // rows x 49  k-DOP extents of P
  DBuf<double> khi;    // rows x 49
  // single-precision filter of the 49-DOP gate of k_narrow (gjk.cuh: kdop_point_gate): per row TOB_KF_ROW floats =
  // 49 x (below threshold, above threshold) relative to the row's centre, then the row's error allowance; kc = the centre
  DBuf<float> kf;      // rows x TOB_KF_ROW
  DBuf<double> kc;     // rows x 4 (centre xyz, unused)
};

struct Level {         // one level of the 32-wide LBVH, SoA boxes
  uint32_t count = 0;
  double* lo[3] = {nullptr, nullptr, nullptr};
  double* hi[3] = {nullptr, nullptr, nullptr};
};

void comm_release(tob_ctx* c);   // comm.cu: destroys an owned NCCL communicator

}  // namespace tob

struct tob_ctx {
  int device = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  tob_params prm{};
  bool have_params = false, have_tables = false;
  int n_tr = 0, T = 0;

  // tables: host copies + device
  std::vector<double> h_basis, h_weight, h_convert, h_mdyn, h_kdop;
  tob::DBuf<double> d_basis, d_weight, d_convert, d_mdyn, d_kdop;
  tob::DBuf<double> d_steps;   // 0.8^k ladder, built by repeated multiplication like the reference's step*=0.8

  // cloud + LBVH
  uint32_t n_pts = 0, n_pad = 0;
  tob::DBuf<double> px, py, pz;
  tob::DBuf<uint32_t> pid;
  std::vector<uint32_t> h_pid;
  tob::DBuf<double> lvl_store;
  // batched independent problems: one cloud per robot slot (empty vectors = one cloud shared by all robots)
  std::vector<uint32_t> cloud_n1, cloud_l1;       // level-1 node count / first level-1 node of each cloud
  std::vector<uint32_t> h_row_task;
  tob::DBuf<uint32_t> row_task, row_l1;           // rows+1: exclusive prefix of broadphase tasks per row; first level-1 node
  tob::DBuf<uint32_t> cta_row;                    // row of the first task of every CTA of a whole-context query
  uint32_t cta_row_tpc = 0;                       // ... for this many tasks per CTA
  int n_levels = 0;
  tob::Level lvl[TOB_MAX_LEVELS];

  // robot states on device; all arrays hold prm.uav_num robots.  Owned robots = [own_begin, own_end)
  int own_begin = 0, own_end = 0;
  tob_allgather_fn ag = nullptr;      // legacy exchange through host callbacks (tob_set_shard); FP64 payloads only
  tob_allreduce_fn ar = nullptr;
  void* cb_user = nullptr;
  // native NCCL exchange (comm.cu): communicator over the ranks that share the robots of this context
  void* nccl_comm = nullptr;          // ncclComm_t
  bool nccl_owned = false;
  int comm_rank = 0, comm_world = 1;
  std::vector<int> shard_first, shard_count;   // robot range of every rank (block partition)
  tob::DBuf<double> ovf_all;          // one word per rank: its overflow / error bits of the current iteration
  tob::DBuf<double> cpl_part, cpl_zy; // coupled solve: 7 Schur sums per robot (exchanged) / z, y of the owned robots
  bool sharded() const { return ag != nullptr || nccl_comm != nullptr; }
  // the NCCL exchanges of an iteration run on a side stream forked from / joined to `stream` (also inside the captured
  // graph), next to the obstacle work of the owned robots that does not need the other robots' data
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t xs = nullptr;          // stream the exchange functions use (null: `stream`)
  bool states_valid = false;
  tob::DBuf<double> s_spline, s_ptime, s_pslack, s_tslack, s_plambda, s_tlambda;
  tob::DBuf<double> s_dir, s_tdir, s_wolfe, s_gnorm;
  tob::DBuf<double> s_step, s_selfstep, s_ptrial, s_e0, s_e1;
  tob::DBuf<unsigned> s_lsmask;       // per robot: rungs 0..31 at which a velocity / acceleration bound is violated for certain
  tob::DBuf<double> s_tstep, s_ttime, s_etr;   // batched line-search trials: robots x TOB_LS_TRIALS
  tob::DBuf<int> s_done, solve_status;

  // per-row geometry + broadphase scratch
  tob::RowGeom geo;
  tob::DBuf<uint32_t> task_cnt, task_off, scan_tmp;
  tob::DBuf<uint32_t> bsum;           // per-CTA totals of the broadphase count pass, scanned in place
  tob::DBuf<uint32_t> cand_pt, cand_row;
  // what the count pass found per item (hit task x hit leaf), so that the fill pass scatters without testing again:
  // row, leaf and point mask of item i of a CTA at rec_*[cta_ibase[cta] + i]; row_li = local item index where a row starts
  tob::DBuf<uint32_t> rec_row, rec_leaf, rec_pm, cta_ibase, cta_nitems, row_li;
  uint64_t rec_cap = 0;
  tob::DBuf<uint32_t> row_off;        // (all rows)+1 candidate offsets; rows outside the queried range are empty
  uint64_t cand_cap = 0;              // capacity (candidates) of cand_pt / cand_row / cpl / cflag; planes: cand_cap + self
  uint64_t n_cand = 0;                // host mirror, valid after sync_counts()
  tob::DBuf<tob::DevCounts> dc;       // device counts (1 element)
  tob::DevCounts* h_dc = nullptr;     // pinned mirror

  // planes: candidate-indexed scratch, then packed CSR over ALL rows
  tob::DBuf<double> cpl;              // cand x 4 (cx,cy,cz,d)
  tob::DBuf<uint32_t> cflag, cflag_off;
  tob::DBuf<uint32_t> csum;           // accepted planes per 128-candidate chunk, scanned in place
  tob::DBuf<uint32_t> selfpre;        // rows+1: exclusive scan of the inter-robot plane count per row
  tob::DBuf<uint32_t> selfcnt;        // rows: inter-robot planes per row (integer atomics of k_self_planes)
  tob::DBuf<double> pl;               // planes x 4
  tob::DBuf<uint32_t> pl_row, pl_off; // plane -> row ; rows+1 offsets
  tob::DBuf<uint32_t> row_nob, row_ntot;
  uint64_t n_planes = 0;

  // persistent planes ("optimal_plane": 1; the reference's is_seperate / seperate_c / seperate_d, Main/admmPathPlanning3D.cpp
  // :342-351, as a sorted sparse set instead of dense N_tr x N_pts arrays): key = row << 32 | index into the sorted cloud
  tob::DBuf<unsigned long long> live_key, live_key_t, new_key;
  tob::DBuf<double> live_pl, live_pl_t, new_pl;   // x 4 (cx, cy, cz, d)
  uint64_t live_cap = 0;
  tob::DBuf<uint32_t> self_live;      // n_tr x npairs: is_self_seperate (Main/multiPathPlanning3D.cpp:450-464)
  tob::DBuf<double> self_lpl;         // n_tr x npairs x 4: self_seperate_c / _d

  // inter-robot scratch
  tob::DBuf<double> self_pl;          // n_tr x npairs x 4
  tob::DBuf<uint32_t> self_ok;        // n_tr x npairs
  tob::DBuf<uint32_t> self_hits;      // inter-robot CCD: colliding (slot, pair) ids [cap] + count [2] + the sorted list [cap]; cap = all tasks

  // energy / gradient / solve scratch
  tob::DBuf<uint32_t> en_items;       // virtual warps v >= 1 of the heavy rows of the current plane set (row << 3 | v - 1)
  tob::DBuf<uint32_t> en_item_base;   // rows: position of the row's first item
  tob::DBuf<double> gpart;            // items x 54: partial plane sums of the gradient pass
  tob::DBuf<double> row_e;            // trials x rows x TOB_EN_REC
  tob::DBuf<int> row_bad;             // robots x TOB_LS_TRIALS: trial infeasible (some d <= 0)
  tob::DBuf<double> row_terms;
  tob::DBuf<double> pc_g, pc_h;
  tob::DBuf<int> pc_flag;
  tob::DBuf<double> band;
  tob::DBuf<int> kmax;
  tob::DBuf<double> red;              // small device scratch (>= 256 doubles)
  tob::DBuf<double> scratch, scratch2;
  tob::DBuf<uint8_t> scratch8;

  double* h_pinned = nullptr;         // pinned read-back area (>= 64 KiB, grown with uav_num: holds one double per robot)
  size_t h_pinned_bytes = 0;
  double* h_stage = nullptr;          // pinned staging for packed state transfers
  size_t h_stage_cap = 0;
  // LBVH build: device time and points of the last tob_cloud_upload[_batch] (roofline entry of the build, 128 B / point)
  double build_ms = 0;
  uint64_t build_points = 0;

  tob_counters ctr{};
  bool bcr_attr_set = false;          // cudaFuncSetAttribute(k_solve_bcr) done on this device

  // whole-iteration CUDA graph (single context, no callbacks, profiling off)
  cudaGraphExec_t graph_exec = nullptr;
  unsigned long long graph_gen = 0;   // alloc_generation() the graph was captured at
  unsigned long long graph_warm_gen = ~0ull;   // generation at which a plain (uncaptured) iteration last ran allocation-free
  int graph_mode = -1;
  uint64_t graph_nodes = 0;           // kernel nodes per launch (for the launch counter)
  bool use_graph = true;
  bool capturing = false;
  // line-search policy of the current iteration: trial points evaluated in the first round / in later rounds (index 0 =
  // current point) and rounds launched ahead.  Few rows: 9, 9, 2 (one launch covers 8 rungs: latency).  Many rows: 3, 9, 3:
  // most robots accept one of the first two rungs, so the first round only evaluates those for everybody (throughput),
  // and the few robots that keep backtracking get 8 rungs per later round
  int ls_kte0 = TOB_LS_TRIALS, ls_kte = TOB_LS_TRIALS, ls_rounds = 2;
  int ls_kte1 = TOB_LS_TRIALS;        // trial slots of round 1 (round 0: ls_kte0, rounds 2..: ls_kte)
  bool ls_e0_ready = false;           // trial slot 0 (the current point) already holds its energy: written by the gradient pass

  // optional per-kernel CUDA-event timing (bench.py roofline): off by default
  bool prof_on = false;
  struct ProfRec { int kid; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[32] = {0};
  uint64_t prof_n[32] = {0};

  // persistent OBSTACLE planes exist only on the single-UAV path (Optimization3D_admm::separate_plane :126-193; the multi-UAV
  // separate_plane, Optimization3D_multi.h:176-235, has no such branch): one robot, or independent problems
  tob_ctx() = default;
  tob_ctx(const tob_ctx&) = delete;
  tob_ctx& operator=(const tob_ctx&) = delete;
  ~tob_ctx() {                        // device buffers release themselves (DBuf); the rest is owned here
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    tob::comm_release(this);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    for (auto& r : prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : prof_pool) cudaEventDestroy(e);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (h_stage) cudaFreeHost(h_stage);
    if (h_dc) cudaFreeHost(h_dc);
    if (stream) cudaStreamDestroy(stream);
  }
  bool live_planes() const { return prm.optimal_plane != 0 && (prm.uav_num == 1 || !cloud_n1.empty()); }
  int n_robots() const { return prm.uav_num; }
  int rows_all() const { return prm.uav_num * n_tr; }
};

namespace tob {

int fail(tob_ctx* c, const char* what, cudaError_t e, const char* file, int line);
int fail_msg(tob_ctx* c, const std::string& msg);

#define TOB_CUDA(c, call)                                                          \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) return tob::fail((c), #call, e__, __FILE__, __LINE__); \
  } while (0)
#define TOB_TRY(expr)    \
  do {                   \
    int r__ = (expr);    \
    if (r__) return r__; \
  } while (0)
#define TOB_LAUNCH_CHECK(c)                                                                  \
  do {                                                                                       \
    (c)->ctr.kernel_launches++;                                                              \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) return tob::fail((c), "kernel launch", e__, __FILE__, __LINE__); \
  } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// the state-changing tail of an iteration (apply step, slack/dual update) runs only if nothing overflowed and every
// robot finished its Armijo search in the rounds that were launched ahead
__device__ __forceinline__ bool iteration_blocked(const DevCounts* dc) {
  return dc != nullptr && (dc->overflow != 0u || dc->ls_pending[dc->ls_rounds - 1] > 0);
}
#endif

// kernel ids of the per-kernel timing (names in api.cu: kKernelNames)
enum KernelId {
  K_ROWS = 0, K_BP_COUNT, K_BP_FILL, K_SCAN, K_NARROW, K_PACK, K_SELF_PLANES, K_ROW_ENERGY, K_ROBOT_ENERGY, K_ROW_GRAD,
  K_PIECE, K_SOLVE, K_CCD, K_SELF_CCD, K_SLACK, K_LINESEARCH_MISC, K_LIVE_REFINE, K_COUNT
};

// scoped CUDA-event pair around ONE kernel launch on the context's stream (no-op unless profiling is enabled)
struct Prof {
  tob_ctx* c;
  bool on;
  Prof(tob_ctx* ctx, int kid) : c(ctx), on(ctx->prof_on) {
    if (!on) return;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; i++) {
      if (!c->prof_pool.empty()) { ev[i] = c->prof_pool.back(); c->prof_pool.pop_back(); }
      else cudaEventCreate(&ev[i]);
    }
    cudaEventRecord(ev[0], c->stream);
    c->prof_pending.push_back({kid, ev[0], ev[1]});
  }
  ~Prof() {
    if (on) cudaEventRecord(c->prof_pending.back().b, c->stream);
  }
};
void prof_collect(tob_ctx* c);   // after a stream sync: fold pending event pairs into prof_ms / prof_n

// ---- module entry points (host side of each .cu) -------------------------------------------------------------
// tables.cu
int make_tables_host(const tob_params& p, const double* time_weight, std::vector<double>& basis, std::vector<double>& weight,
                     std::vector<double>& convert, std::vector<double>& mdyn, std::vector<double>& kdop);
// lbvh.cu
int lbvh_build(tob_ctx* c, const double* V_host, uint32_t n);
int lbvh_build_batch(tob_ctx* c, const double* const* V_host, const uint32_t* n, int n_clouds);
// queries rows [rb*n_tr, re*n_tr) whose boxes are in geo.box; fills cand_pt/cand_row (GLOBAL rows) and row_off.
// Asynchronous: the total stays on the device (dc->n_cand); count_as = 1 adds it to the dcd_candidates counter.
int broadphase(tob_ctx* c, int rb, int re, double d, int count_as);
int broadphase_rows(tob_ctx* c, int row_base, int rows, double d, int count_as);
int ensure_query_buffers(tob_ctx* c);
// api.cu: device counts -> pinned mirror (one stream sync); capacity growth after an overflow
int sync_counts(tob_ctx* c);
int grow_cand_capacity(tob_ctx* c, uint64_t need);
// segments.cu : mode bits: 1 = k-DOP extents, 2 = direction rows + swept box, 4 = trial point spline+step*dir
int compute_rows(tob_ctx* c, const double* spline_dev, const double* dir_dev, const double* step_dev, int rb, int re, int mode);
// narrow.cu
int narrowphase_planes(tob_ctx* c, int rb, int re, int with_self);   // with_self < 0: obstacle candidates only, no packing
int narrowphase_finish(tob_ctx* c, int rb, int re, int with_self);
int pack_planes_from_host(tob_ctx* c, int rb, int re, const uint32_t* offsets, const double* cc, const double* dd);
int ccd_position_steps(tob_ctx* c, int rb, int re);
int self_planes(tob_ctx* c);
int pack_self_only(tob_ctx* c);
int ensure_live_buffers(tob_ctx* c, uint64_t need);     // persistent-plane set: grows preserving the live planes
int reset_live_planes(tob_ctx* c);
int self_ccd_steps(tob_ctx* c, int coupled, double* steps_dev);
int edge_validity(tob_ctx* c, const double* edges_host, int n, double d, uint8_t* valid_host);
// barrier.cu
int energy_trials(tob_ctx* c, int rb, int re, const double* dir, const double* tstep, const double* ttime, int KT, int k0,
                  int k1, double* e_dev);
// one Armijo round of robots [rb,re) (decoupled): trial energies k0..kte-1 (kte <= TOB_LS_TRIALS) + the ladder decision, robots that
// are already done are skipped on the device; robots still backtracking are counted in dc->ls_pending[slot]
int line_search_round(tob_ctx* c, int rb, int re, int wolfe_idx, int k0, int kte, int slot, int k0e);
int line_search_begin(tob_ctx* c, int rb, int re);
int energy_items(tob_ctx* c, bool reset);   // after every new plane CSR: lists the extra virtual warps of the heavy rows   // clears the infeasibility flags of the robots' trial slots
#define TOB_EN_REC 9   // doubles per (trial, row) in row_e: 8 plane-energy partials + the bound energy (barrier.cu: EN_REC)
int gradient_blocks(tob_ctx* c, int rb, int re, int project_psd);
int row_blocks(tob_ctx* c, int tr, int which, double* out_dev);
// comm.cu: in-place exchange of robot-indexed device arrays between the ranks (no-op when the context is not sharded)
int exchange_robots(tob_ctx* c, void* buf, size_t elems_per_robot, size_t elem_size);
int exchange_ranks(tob_ctx* c, double* buf);    // one FP64 word per rank
int exchange_group_begin(tob_ctx* c);
int exchange_group_end(tob_ctx* c);
int exchange_fork(tob_ctx* c);    // the exchanges that follow run on the side stream, after what is on the main stream now
int exchange_join(tob_ctx* c);    // the main stream waits for them
void shard_partition(tob_ctx* c);
// solve.cu
int solve_directions(tob_ctx* c, int rb, int re, int dense_shift);
int solve_coupled(tob_ctx* c);
// guarded != 0: no-op when the iteration cannot be committed (overflow / line search unfinished); commits otherwise
int slack_update(tob_ctx* c, int rb, int re, int guarded = 0);
int slack_terms(tob_ctx* c, const double* in57_dev, int consensus, double* out381_dev);

}  // namespace tob
