/* trajopt_b200.h -- C ABI of the B200-native hot path of traj-opt-admm.
 *
 * The reference has no FFI layer: its hot path sits behind header-only C++ statics (namespace HighOrderCCD) and
 * the compiled class BVH.  This C ABI is the thin layer those entry points are re-hosted on: every function
 * below names the reference interface (file:line under the reference tree) it replaces.  The C++ drop-in
 * headers in traj-opt-admm_b200/host/HighOrderCCD/ keep the reference signatures and marshal Eigen objects into
 * these calls; the ctypes binding in traj-opt-admm_b200/trajopt/api.py does the same for Python.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types.  All pointers are HOST pointers unless a name ends in _dev.
 *   - matrices are column-major FP64 exactly like Eigen::MatrixXd (spline: T x 3, p_slack/p_lambda: 6P x 3).
 *   - point / segment ids are uint32 (BVH.h:16 uses unsigned int).
 *   - "row" r = robot*n_tr + tr_id addresses one Bezier sub-segment of one robot; ragged per-row lists are CSR:
 *     offsets[n_rows+1] + payload.
 *   - every function returns 0 on success, nonzero on failure (tob_last_error() gives the text).  There is no
 *     CPU fallback: a missing GPU or a CUDA error is an error.
 *   - in-band numerical conventions of the reference are kept: an infeasible trial point yields +INFINITY
 *     energy (Energy_admm.h:81-82,137-138,154-155).
 */
#ifndef TRAJOPT_B200_H
#define TRAJOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tob_ctx tob_ctx;

/* Solver globals of HighOrderCCD/Utils/CCDUtils.cpp:5-44 that the path reads (set from Config_File/3D.json in
 * Main/admmPathPlanning3D.cpp:368-397 and hard-coded :477-478 / multiPathPlanning3D.cpp:596). */
typedef struct tob_params {
  int32_t piece_num;     /* Bezier pieces per robot (order 5, 6 control points each) */
  int32_t res;           /* sub-segments per piece; n_tr = piece_num*res */
  int32_t uav_num;       /* robots resident in this context */
  int32_t optimal_plane; /* is_optimal_plane: 0 (3D.json) = planes rebuilt every iteration; 1 = persistent planes refined by
                          * Optimal_plane::optimal_cd / self_optimal_cd (see tob_planes_reset) */
  double lambda;         /* barrier weight */
  double margin;         /* barrier activation distance d-hat */
  double offset;         /* safety distance */
  double mu;             /* ADMM penalty */
  double vel_limit, acc_limit;
  double ks, kt;         /* smoothness / time weights */
} tob_params;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
int tob_ctx_create(int device, tob_ctx** out);
void tob_ctx_destroy(tob_ctx* ctx);
const char* tob_last_error(const tob_ctx* ctx);          /* ctx may be NULL: last error of tob_ctx_create */
int tob_device_info(const tob_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor);

/* ---- set-up (replaces the set-up code of Main/admmPathPlanning3D.cpp:403-414,448-468,294-338) -------------- */
int tob_set_params(tob_ctx* ctx, const tob_params* p);
/* Computes on the host, with the reference's formulas and evaluation order, and uploads: Combination<40>
 * (CCDUtils.h:110-138), Conversion<5> (:140-170), Dynamic3D<5,3> (:172-226), Blossom<5> x convert (:228-315 and
 * admmPathPlanning3D.cpp:303-313), normalised k-DOP axes (CCDUtils.cpp:56-119).  time_weight may be NULL (=1). */
int tob_make_tables(tob_ctx* ctx, const double* time_weight);
/* Upload tables computed elsewhere (the C++ drop-in passes the reference globals subdivide_tree, convert_list,
 * M_dynamic, kdop_matrix).  basis: n_tr x 36 (6x6 col-major each), weight: n_tr, convert: piece_num x 36,
 * mdyn: 36, kdop: 3x49 col-major. */
int tob_set_tables(tob_ctx* ctx, const double* basis, const double* weight, const double* convert,
                   const double* mdyn, const double* kdop);
int tob_get_tables(const tob_ctx* ctx, double* basis, double* weight, double* convert, double* mdyn, double* kdop);

/* ---- point cloud (replaces BVH::InitPointcloud, BVH/BVH.cpp:53-92) ------------------------------------------ */
/* V: n x 3 column-major.  Builds the Morton-sorted 32-wide LBVH on the device. */
int tob_cloud_upload(tob_ctx* ctx, const double* V, uint32_t n);
/* Batched independent problems (one BVH::InitPointcloud per problem): cloud b (n[b] x 3 column-major at V[b]) belongs to
 * robot slot b, n_clouds == tob_params.uav_num.  The clouds are concatenated on the device (each padded to a multiple of
 * 1024 points) and every row only walks the level-1 nodes of its own cloud.  Use with mode 2 of tob_admm_iterate /
 * tob_optimization; call after tob_set_params.  Point ids returned by the broadphase entry points are cloud-local. */
int tob_cloud_upload_batch(tob_ctx* ctx, const double* const* V, const uint32_t* n, int n_clouds);
uint32_t tob_cloud_size(const tob_ctx* ctx);

/* ---- broadphase (replaces BVH::DCDCollision :149-192, BVH::CCDCollision :195-249) --------------------------- */
/* splines: n_robots blocks of T x 3.  Output CSR over n_robots*n_tr rows; ids are ORIGINAL point ids, sorted
 * ascending inside each row (the reference returns its tree's DFS order; the contract is the SET, see
 * BVH/src/AABB.cc:131-161 for the leaf predicate that defines it).  If total > cap nothing is written to ids. */
int tob_broadphase_dcd(tob_ctx* ctx, const double* splines, int n_robots, double d, uint32_t* offsets,
                       uint32_t* ids, uint64_t cap, uint64_t* total);
int tob_broadphase_ccd(tob_ctx* ctx, const double* splines, const double* directions, int n_robots, double d,
                       uint32_t* offsets, uint32_t* ids, uint64_t cap, uint64_t* total);
/* replaces BVH::SelfDCDCollision :252-286 / SelfCCDCollision :289-329 for one time slot: P (and D): u blocks of
 * 6x3 col-major.  pairs: (first<second) sorted lexicographically. */
int tob_self_broadphase(tob_ctx* ctx, const double* P, const double* D /* NULL = DCD */, int u, double d,
                        uint32_t* pairs, uint64_t cap, uint64_t* total);

/* replaces BVH::EdgeCollision (BVH/BVH.cpp:95-133, front-end path validation): all points within d of the box
 * [lo,hi] (leaf predicate of AABB.cc:131-161); ids = original point ids, ascending. */
int tob_box_query(tob_ctx* ctx, const double* lo, const double* hi, double d, uint32_t* ids, uint64_t cap,
                  uint64_t* total);

/* The obstacle loop of the front end's motion validator (HighOrderCCD/OMPL/OMPL.cpp:36-96: BVH::EdgeCollision, then
 * CCD::GJKDCD(edge, point, d) per candidate, CCD/CCD.h:17-114; invalid at the first collision), batched: edges = n x 2 x 3
 * (edge, endpoint, xyz); valid[i] = 1 iff no cloud point is within d of the segment i.  One launch for all edges. */
int tob_edge_validity_batch(tob_ctx* ctx, const double* edges, int n, double d, uint8_t* valid);

/* ---- per-pair primitives (function-level parity entry points) ----------------------------------------------- */
/* batch of n independent evaluations of gjk() (lib/opengjk/src/openGJK.c:754-852) as marshalled by CCD::GJKDCD /
 * GJKCCD / SelfGJKCCD (CCD/CCD.h:17-352); A: n blocks of na x 3 col-major, B likewise (1 <= na,nb <= 16). */
int tob_gjk_batch(tob_ctx* ctx, const double* A, int na, const double* B, int nb, int n, double* v /* n x 3 */);
/* 49-DOP separation test of two vertex sets with gap d: CCD::KDOPDCD :354-413 (6,1), SelfKDOPDCD :535-587 (6,6),
 * KDOPCCD :416-473 (12,1), SelfKDOPCCD :475-533 (12,12); the swept sets [P+tMin*D; P+tMax*D] are formed by the caller */
int tob_kdop_batch(tob_ctx* ctx, const double* A, int na, const double* B, int nb, int n, double d, uint8_t* flags);
/* CCD::KDOPDCD (CCD/CCD.h:354-413): P n x (6x3), q n x 3 -> flags */
int tob_kdop_dcd_batch(tob_ctx* ctx, const double* P, const double* q, int n, double d, uint8_t* flags);
/* Separate::opengjk (Separate.h:18-163): -> ok flag, c (n x 3), d (n) */
int tob_plane_point_batch(tob_ctx* ctx, const double* P, const double* q, int n, double distance, uint8_t* ok,
                          double* c, double* d);
/* Separate::selfgjk (:165-304) followed (if refine != 0) by Optimal_plane::optimal_d (Optimal_plane.h:13-71) */
int tob_plane_hulls_batch(tob_ctx* ctx, const double* P0, const double* P1, int n, double distance, int refine,
                          uint8_t* ok, double* c, double* d);

/* Optimal_plane::optimal_d (Optimal_plane.h:13-71): 1-D Newton on d for n (P0, P1, c) triples; d_io in/out */
int tob_refine_d_batch(tob_ctx* ctx, const double* P0, const double* P1, const double* c, int n, double* d_io);

/* Optimal_plane::optimal_cd (Optimal_plane.h:160-293): Newton on the tangent angles of c for n (sub-segment P 6x3
 * column-major, obstacle point q) pairs, d = -c.q - offset; c (n x 3) and d_io (n) in/out.  capped[i] != 0 (may be NULL):
 * pair i left through a loop cap instead of one of the reference's two exits (its loops are unbounded). */
int tob_optimal_cd_batch(tob_ctx* ctx, const double* P, const double* q, int n, double* c, double* d_io, uint8_t* capped);
/* Optimal_plane::self_optimal_cd (Optimal_plane.h:620-773): (theta, phi, d) Newton for n inter-robot pairs */
int tob_self_optimal_cd_batch(tob_ctx* ctx, const double* P0, const double* P1, int n, double* c, double* d_io, uint8_t* capped);

/* ---- persistent planes (tob_params.optimal_plane = 1) ---------------------------------------------------------------
 * Replaces the globals is_seperate / seperate_c / seperate_d (CCDUtils.cpp:34-36, sized N_tr x N_pts in
 * Main/admmPathPlanning3D.cpp:342-351) and is_self_seperate / self_seperate_c / _d (CCDUtils.cpp:30-32,
 * Main/multiPathPlanning3D.cpp:450-464).  The obstacle set is a sorted sparse list on the device (single-UAV contexts and
 * batched independent problems; the multi-UAV separate_plane of the reference has no persistent branch), the inter-robot
 * set a dense (slot, pair) table.  Every plane pass (tob_separate_planes, tob_admm_iterate, tob_optimization) adds the
 * planes of pairs that were not live and refines ALL live planes, like Optimization3D_admm.h:126-193 /
 * Optimization3D_multi.h:271-339.  tob_planes_reset empties both sets (also done by tob_set_params and a cloud upload). */
int tob_planes_reset(tob_ctx* ctx);
/* live obstacle planes in (row, position in the sorted cloud) order: rows[k] = robot*n_tr + tr, ids[k] = original point id
 * (cloud-local for batched clouds), c (k x 3), d.  *total = live count; arrays are filled only when total <= cap. */
int tob_live_planes(tob_ctx* ctx, uint32_t* rows, uint32_t* ids, double* c, double* d, uint64_t cap, uint64_t* total);

/* ---- separating planes (replaces Optimization3D_admm::separate_plane, Optimization3D_admm.h:69-197, and
 *      Optimization3D_multi::separate_plane :176-235 / ::separate_self :237-342) ------------------------------ */
/* Runs broadphase -> 49-DOP -> GJK -> plane for every robot and, when with_self != 0 and n_robots > 1, the
 * inter-robot planes appended after the obstacle planes of each row (same order as the reference: obstacle
 * planes first).  The plane set stays resident on the device for the calls below; it is also returned on the
 * host when c/dd are non-NULL.  c: total x 3 (row-major xyz per plane), dd: total. */
int tob_separate_planes(tob_ctx* ctx, const double* splines, int n_robots, int with_self, uint32_t* offsets,
                        double* c, double* dd, uint64_t cap, uint64_t* total);
/* Optimization3D_multi::separate_self (:237-342) alone: the inter-robot planes of every (robot, time slot) as a CSR over
 * n_robots*n_tr rows, (c, d-offset/2) for the lower robot id of a pair and (-c, -d-offset/2) for the higher one.  They
 * also become the resident plane set. */
int tob_separate_self(tob_ctx* ctx, const double* splines, int n_robots, uint32_t* offsets, double* c, double* dd,
                      uint64_t cap, uint64_t* total);
/* Upload a caller-provided plane set (the reference's c_lists/d_lists) as the resident set. */
int tob_set_planes(tob_ctx* ctx, int n_robots, const uint32_t* offsets, const double* c, const double* dd);

/* ---- energies (replace Energy_admm::*, Energy_admm.h) ------------------------------------------------------- */
typedef struct tob_state {       /* one robot's ADMM variables, host pointers */
  double* spline;                /* T x 3 */
  double* piece_time;            /* scalar */
  double* p_slack;               /* 6P x 3 */
  double* t_slack;               /* P */
  double* p_lambda;              /* 6P x 3 */
  double* t_lambda;              /* P */
} tob_state;

/* Energy_admm::plane_barrier_energy :46-96 against the resident planes of `robot` */
int tob_plane_barrier_energy(tob_ctx* ctx, int robot, const double* spline, double* e);
/* Energy_admm::bound_energy :98-170 */
int tob_bound_energy(tob_ctx* ctx, const double* spline, double piece_time, double* e);
/* Energy_admm::spline_energy :16-44 */
int tob_spline_energy(tob_ctx* ctx, int robot, const tob_state* st, double* e);

/* ---- gradient / Hessian blocks (replace Gradient_admm::*, Gradient_admm.h) ---------------------------------- */
/* local_spline_gradient :67-164 for every piece (before the PSD projection): g: P x 19, h: P x 361 (col-major) */
int tob_piece_blocks(tob_ctx* ctx, int robot, const tob_state* st, int project_psd, double* g, double* h);
/* global_spline_gradient :13-65: dense grad[3T+1], hess[(3T+1)^2 col-major] (PSD-projected piece blocks) */
int tob_global_gradient(tob_ctx* ctx, int robot, const tob_state* st, double* grad, double* hess);
/* Optimization3D_admm::spline_descent_direction (Optimization3D_admm.h:400-503) when dense_shift == 0;
 * Optimization3D_multi::spline_descent_direction (Optimization3D_multi.h:659-752, global eigen-shift fall-back)
 * when dense_shift != 0.  direction: T x 3. */
int tob_descent_direction(tob_ctx* ctx, int robot, const tob_state* st, int dense_shift, double* direction,
                          double* t_direction, double* wolfe, double* gnorm);

/* One sub-segment (row tr_id of robot 0) against the resident planes, in the 18 piece coordinates (control point m, axis k
 * at index 3m+k), without the lambda factor: which = 0 -> Gradient_admm::local_plane_barrier_gradient (Gradient_admm.h:331-407);
 * which = 1 -> local_bound_gradient (:409-572) incl. g_t, h_t and the mixed column partgrad.  hess324: 18x18 col-major. */
int tob_row_blocks(tob_ctx* ctx, const double* spline, double piece_time, int tr_id, int which, double* grad18,
                   double* hess324, double* g_t, double* h_t, double* partgrad18);

/* ---- line search (replaces Optimization3D_admm::spline_line_search, Optimization3D_admm.h:505-557, when *step_io < 0:
 *      the bound is Step::position_step; and Optimization3D_multi::spline_line_search :754-811 when 0 <= *step_io <= 1:
 *      the caller's bound).  Armijo backtracking with factor 0.8 against the resident planes; st->spline and
 *      st->piece_time are updated in place, *step_io receives the accepted step. */
int tob_line_search(tob_ctx* ctx, int robot, tob_state* st, const double* direction, double t_direction, double wolfe,
                    double* step_io);
/* the reference's global `wolfe` as left by the direction solve of the last iteration (Optimization3D_admm.h:477) */
int tob_last_wolfe(tob_ctx* ctx, double* wolfe);

/* ---- CCD step bound (replaces Step::position_step Step.h:21-110, ::self_step :184-256,
 *      ::couple_self_step :112-182) ---------------------------------------------------------------------------- */
int tob_position_step(tob_ctx* ctx, const double* spline, const double* direction, double* step);
int tob_self_step(tob_ctx* ctx, const double* splines, const double* directions, int n_robots, int coupled,
                  double* steps /* n_robots, or 1 when coupled */);

/* ---- slack / dual update (replaces Optimization3D_admm::update_slack_lambda :231-398) ------------------------ */
int tob_update_slack_lambda(tob_ctx* ctx, tob_state* st);

/* One piece of the slack problem.  consensus != 0: Energy_admm::slack_energy (Energy_admm.h:172-190) and
 * Gradient_admm::slack_gradient (Gradient_admm.h:574-622).  consensus == 0: Energy_admm::dynamic_energy (:199-215) and
 * Gradient_admm::dynamic_gradient (:633-671); then c_spline / p_lambda may be NULL, grad19[18] = g_t,
 * hess361(18,18) = h_t and hess361(0..17,18) = partgrad.  c_spline, p_part, p_lambda: 6x3 column-major.
 * energy / grad19 / hess361 (19x19 column-major) may each be NULL. */
int tob_slack_terms(tob_ctx* ctx, const double* c_spline, double piece_time, const double* p_part, double t_part,
                    const double* p_lambda, double t_lambda, int consensus, double* energy, double* grad19,
                    double* hess361);

/* ---- whole ADMM iteration, device resident --------------------------------------------------------------------
 * tob_states_upload / download move the n_robots states between host and the context;
 * tob_admm_iterate runs `iters` iterations of Optimization3D_admm::optimization (:29-67) when n_robots == 1, of
 * Optimization3D_multi::optimization_decouple (Optimization3D_multi.h:29-118) when n_robots > 1 (mode 0) or of
 * ::optimization (coupled, :120-174) (mode 1); mode 2 treats the n_robots slots as INDEPENDENT single-UAV problems (each one
 * iterates like Optimization3D_admm::optimization against its own cloud, no inter-robot terms; gnorm = mean over the
 * problems).  gnorm receives the reference's global `gnorm` after the last iteration.  Mode 1 shares ONE piece time: states[u].piece_time may all point to the same double.  Robot ownership for
 * multi-GPU: see tob_set_shard() (mode 0 only). */
int tob_states_upload(tob_ctx* ctx, const tob_state* states, int n_robots);
int tob_states_download(tob_ctx* ctx, tob_state* states, int n_robots);
int tob_admm_iterate(tob_ctx* ctx, int iters, int mode, double* gnorm);
/* one call = upload + 1 iteration + download: the shape of the reference entry point (host in / host out) */
int tob_optimization(tob_ctx* ctx, tob_state* states, int n_robots, int mode, double* gnorm);

/* counters of the last tob_admm_iterate / tob_optimization call (for bench.py): */
typedef struct tob_counters {
  uint64_t kernel_launches;      /* kernels of this library launched */
  uint64_t dcd_candidates;       /* broadphase candidates through k-DOP (+GJK) */
  uint64_t planes;               /* accepted planes */
  uint64_t ccd_candidates;       /* swept-box candidates through the CCD ladder */
  uint64_t energy_plane_evals;   /* planes x energy/gradient passes */
  uint64_t self_pairs;           /* inter-robot segment pairs evaluated */
  uint64_t line_search_trials;
  uint64_t barrier_terms;        /* (control point, plane) terms inside the barrier band (d < margin) that were evaluated */
  uint64_t live_planes;          /* persistent-plane mode: live (sub-segment, point) planes */
  uint64_t refine_capped;        /* plane refinements stopped by a loop cap (the reference's loops are unbounded) */
  /* counted work (not a model): what the narrowphase / CCD kernels really executed */
  uint64_t np_kdop_groups;       /* 7-axis groups of the 49-DOP gate evaluated (49 axes = 7 groups, early exit between groups; FP32 filter) */
  uint64_t np_gjk_iters;         /* GJK(6,1) rounds run for the k-DOP survivors */
  uint64_t ccd_gjk_iters;        /* GJK(12,1) rounds run by the CCD ladder */
  uint64_t ccd_kdop_pass;        /* swept candidates that passed the swept 49-DOP gate */
  uint64_t np_kdop_exact;        /* axes of the 49-DOP gate the single-precision filter left undecided (re-tested in FP64) */
  uint64_t np_band;              /* pairs whose GJK distance fell within 1e-6 (relative) of the gap: the only ones that need axes 15..49 of the gate */
  uint64_t ls_rung_hist[8];      /* decoupled line searches by accepted rung of the 0.8 ladder (counted from the first rung that was evaluated): [0] .. [7] = the eighth or deeper */
  uint64_t ls_rungs_skipped;     /* rungs not evaluated because a velocity / acceleration bound is violated for certain at that step (their energy is +inf) */
} tob_counters;
int tob_get_counters(const tob_ctx* ctx, tob_counters* out);
int tob_reset_counters(tob_ctx* ctx);

/* per-kernel device time (CUDA events on the context's stream around each launch of this library's kernels);
 * used by bench.py for the roofline of the dominant kernel.  kid = 0.. until the call returns nonzero. */
int tob_profile_enable(tob_ctx* ctx, int on);
int tob_profile_read(tob_ctx* ctx, int kid, double* ms_total, uint64_t* launches, const char** name);

/* ---- multi-GPU (robots sharded across ranks; cloud replicated) -------------------------------------------------
 * The reference iterates over the robots serially (Optimization3D_multi.h:40-49,59-71,78-89); here every rank (one process
 * or host thread per GPU, one context each) owns a contiguous block of the uav_num robots and the per-iteration exchange
 * runs over NCCL on the context's stream, inside the iteration's CUDA graph:
 *   decoupled (mode 0): all-gather of the control points before separate_self (Optimization3D_multi.h:51), then ONE
 *                       grouped all-gather of directions + wolfe + gnorm before Step::self_step (:72-76)
 *   coupled   (mode 1): additionally the seven Schur sums of every robot's block of the joint Newton system (:519-557),
 *                       the CCD ladder exponents (shared step = min over robots, :586-594) and the trial energies of each
 *                       Armijo round (:605-636); all sums are taken locally in robot order after the all-gather.
 * A sharded run is bitwise equal to the same problem in one context.  NCCL is bound at run time (dlopen of libnccl.so.2,
 * reusing a copy the process has already loaded); every rank calls the same sequence of tob_admm_iterate /
 * tob_optimization.  tob_states_upload takes ALL uav_num states on every rank (only the owned ones are used);
 * tob_states_download returns valid data for the owned robots [first, first+count).
 *
 *   tob_nccl_unique_id   rank 0: 128-byte ncclUniqueId to hand to the other ranks (MPI, a file, torch.distributed ...)
 *   tob_nccl_init_rank   ncclCommInitRank on the context's device; the communicator is owned by the context
 *   tob_nccl_attach      use a communicator the host already has (ncclComm_t; not destroyed with the context)
 *   tob_nccl_init_all    one process driving n GPUs (one context each, one host thread per context while iterating):
 *                        ncclCommInitAll over the contexts' devices
 *   tob_nccl_detach      back to an unsharded context
 *   tob_shard_range      robots owned by this context: first = rank*floor(U/W) + min(rank, U mod W)            */
int tob_nccl_available(int* version /* may be NULL; NCCL_VERSION_CODE of the library that was bound */);
int tob_nccl_unique_id(void* id128);
int tob_nccl_init_rank(tob_ctx* ctx, const void* id128, int rank, int world);
int tob_nccl_attach(tob_ctx* ctx, void* nccl_comm);
int tob_nccl_init_all(tob_ctx** ctxs, int n_ctx);
int tob_nccl_detach(tob_ctx* ctx);
int tob_shard_range(const tob_ctx* ctx, int* first, int* count, int* rank, int* world);

/* Legacy exchange through host callbacks (decoupled mode, equal blocks only, not graph-captured): the context owns robots
 * [first, first+count) of n_total and calls
 *   allgather(dev_ptr_full, elems_per_rank, user): in-place all-gather of FP64 on the context's stream.
 * for every exchange above (allreduce is unused: sums are taken locally after the all-gather). */
typedef int (*tob_allgather_fn)(void* dev_ptr_full, uint64_t elems_per_rank, void* user);
typedef int (*tob_allreduce_fn)(void* dev_ptr, uint64_t n, int op /*0 sum,1 min,2 max*/, void* user);
int tob_set_shard(tob_ctx* ctx, int first, int count, int n_total, tob_allgather_fn ag, tob_allreduce_fn ar,
                  void* user);
void* tob_stream(tob_ctx* ctx); /* cudaStream_t the context launches on */

/* LBVH build of the last tob_cloud_upload / tob_cloud_upload_batch: device time (H2D of the clouds + Morton keys + radix
 * sort + gather, CUDA events on the context's stream) and points; the build streams 128 B per point. */
int tob_build_stats(const tob_ctx* ctx, double* ms, uint64_t* points);

/* FP64 pipe microbenchmark (DFMA chains) used as the roofline denominator of the FP64-bound kernels */
int tob_fp64_peak(tob_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
