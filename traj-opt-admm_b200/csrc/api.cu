// api.cu -- extern "C" entry points of include/trajopt_b200.h and the device-resident ADMM iteration.
//
// Iteration sequencing follows Optimization3D_admm::optimization (Optimization3D_admm.h:29-67) for one robot and
// Optimization3D_multi::optimization_decouple (Optimization3D_multi.h:29-118) for several:
//   separate planes -> per-piece blocks -> Newton direction -> CCD step bound -> Armijo line search -> slack/dual.
// Everything stays on the device; the host only reads a few scalars (candidate/plane totals for buffer sizing and
// the number of robots still backtracking).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ctx.cuh"
#include "gjk.cuh"
#include "optplane.cuh"

static std::string g_create_err;

namespace tob {

int fail(tob_ctx* c, const char* what, cudaError_t e, const char* file, int line) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")";
  if (c) c->err = m; else g_create_err = m;
  return 1;
}
int fail_msg(tob_ctx* c, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return 1;
}

void prof_collect(tob_ctx* c) {
  for (auto& r : c->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.kid] += ms; c->prof_n[r.kid]++; }
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
}

static int need(tob_ctx* c, bool tables, bool cloud) {
  if (!c) return fail_msg(c, "null context");
  if (!c->have_params) return fail_msg(c, "tob_set_params has not been called");
  if (tables && !c->have_tables) return fail_msg(c, "tables missing: call tob_make_tables or tob_set_tables");
  if (cloud && c->n_pts == 0) return fail_msg(c, "no point cloud: call tob_cloud_upload");
  return 0;
}

static int upload(tob_ctx* c, DBuf<double>& b, const double* src, size_t n, size_t elem_off = 0) {
  TOB_CUDA(c, cudaMemcpyAsync(b.p + elem_off, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return 0;
}

static int pinned_ensure(tob_ctx* c, size_t bytes);

static int alloc_states(tob_ctx* c) {
  const int U = c->n_robots(), P = c->prm.piece_num, T = c->T;
  TOB_CUDA(c, c->s_spline.ensure((size_t)U * 3 * T));
  TOB_CUDA(c, c->s_ptime.ensure(U));
  TOB_CUDA(c, c->s_pslack.ensure((size_t)U * 18 * P));
  TOB_CUDA(c, c->s_tslack.ensure((size_t)U * P));
  TOB_CUDA(c, c->s_plambda.ensure((size_t)U * 18 * P));
  TOB_CUDA(c, c->s_tlambda.ensure((size_t)U * P));
  TOB_CUDA(c, c->s_dir.ensure((size_t)U * 3 * T));
  TOB_CUDA(c, c->s_tdir.ensure(U)); TOB_CUDA(c, c->s_wolfe.ensure(U)); TOB_CUDA(c, c->s_gnorm.ensure(U));
  TOB_CUDA(c, c->s_step.ensure(U)); TOB_CUDA(c, c->s_selfstep.ensure(U)); TOB_CUDA(c, c->s_ptrial.ensure(U));
  TOB_CUDA(c, c->s_lsmask.ensure(U));
  TOB_CUDA(c, c->s_e0.ensure(U)); TOB_CUDA(c, c->s_e1.ensure(U));
  TOB_CUDA(c, c->s_tstep.ensure((size_t)U * TOB_LS_TRIALS)); TOB_CUDA(c, c->s_ttime.ensure((size_t)U * TOB_LS_TRIALS));
  TOB_CUDA(c, c->s_etr.ensure((size_t)U * TOB_LS_TRIALS));
  TOB_CUDA(c, c->s_done.ensure(U + 1)); TOB_CUDA(c, c->solve_status.ensure(U));
  TOB_CUDA(c, c->kmax.ensure(U + 1));
  TOB_TRY(pinned_ensure(c, (size_t)U * sizeof(double) + 64));   // gnorm of every robot is read back once per iteration
  return 0;
}

static int put_state(tob_ctx* c, int robot, const tob_state* st) {
  const int P = c->prm.piece_num, T = c->T;
  TOB_TRY(upload(c, c->s_spline, st->spline, 3 * T, (size_t)robot * 3 * T));
  TOB_TRY(upload(c, c->s_ptime, st->piece_time, 1, robot));
  TOB_TRY(upload(c, c->s_pslack, st->p_slack, 18 * P, (size_t)robot * 18 * P));
  TOB_TRY(upload(c, c->s_tslack, st->t_slack, P, (size_t)robot * P));
  TOB_TRY(upload(c, c->s_plambda, st->p_lambda, 18 * P, (size_t)robot * 18 * P));
  TOB_TRY(upload(c, c->s_tlambda, st->t_lambda, P, (size_t)robot * P));
  return 0;
}

static int get_state(tob_ctx* c, int robot, tob_state* st) {
  const int P = c->prm.piece_num, T = c->T;
  cudaStream_t s = c->stream;
  TOB_CUDA(c, cudaMemcpyAsync(st->spline, c->s_spline.p + (size_t)robot * 3 * T, 3 * T * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->piece_time, c->s_ptime.p + robot, sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->p_slack, c->s_pslack.p + (size_t)robot * 18 * P, 18 * P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->t_slack, c->s_tslack.p + (size_t)robot * P, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->p_lambda, c->s_plambda.p + (size_t)robot * 18 * P, 18 * P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->t_lambda, c->s_tlambda.p + (size_t)robot * P, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}

// ---- small device helpers of the iteration ---------------------------------------------------------------------
__global__ void k_fill_int(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// first step of the ladder: min(self step, 0.8^kmax), clamped so the piece time stays positive (Optimization3D_admm.h:521-524)
__device__ __forceinline__ double ls_base_step(int u, const int* kmax, const double* steps_tab, const double* selfstep, int use_self,
                                               const double* ptime, const double* tdir) {
  double s = steps_tab[kmax[u]];
  if (use_self) {
    double ss = selfstep[u];
    if (s < ss) ss = s;        // reference: if(step<step_list[i]) step_list[i]=step
    s = ss;
  }
  if (ptime[u] + s * tdir[u] <= 0) s = -0.95 * ptime[u] / tdir[u];
  return s;
}

// Rungs of the 0.8 ladder at which a velocity / acceleration bound of some sub-segment is violated FOR CERTAIN: the
// reference's energy is +inf there (Energy_admm.h:98-170 returns at the first d <= 0) and its Armijo loop moves on
// (Optimization3D_admm.h:537-544).  On the batched workload the step the CCD bound allows shortens the piece time so much
// that the first 8 to 12 rungs violate the velocity bound: evaluated, each costs a pass of k_row_energy up to the first
// warp that meets the violation, and the round that finally is feasible evaluates four rungs of which the first is accepted.
// One thread per (row, rung) of the first 32 rungs evaluates the row's nine bound terms -- trial control points
// basis (x + s dir) and trial time like k_row_energy -- and sets the rung's bit when some d < -1e-9 (limit + 1): six orders of
// magnitude above what the evaluation order can change.  k_ls_init starts the ladder behind the leading set bits; every
// rung from there on is evaluated as before, so a rung is skipped only when its energy is known to be +inf.
struct LsMaskArgs {
  int rb, re, n_tr, res, T, use_self;
  const int* kmax;
  const double *steps_tab, *selfstep, *ptime, *tdir, *spline, *dir, *basis, *weight;
  double vel_limit, acc_limit;
  unsigned* mask;       // per robot
};
__global__ void __launch_bounds__(128) k_ls_bound_mask(LsMaskArgs a) {
  const int rung = threadIdx.x & 31;
  const int ri = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (ri >= (a.re - a.rb) * a.n_tr) return;
  const int u = a.rb + ri / a.n_tr, tr = ri % a.n_tr;
  double s = ls_base_step(u, a.kmax, a.steps_tab, a.selfstep, a.use_self, a.ptime, a.tdir);
  for (int k = 0; k < rung; k++) s *= 0.8;
  const double t = a.ptime[u] + s * a.tdir[u];
  const double w = a.weight[tr];
  const double* B = a.basis + (size_t)36 * tr;
  double P[18];
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    double bz[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const size_t gi = (size_t)u * 3 * a.T + (size_t)ax * a.T + 3 * (tr / a.res) + i;
      bz[i] = a.spline[gi] + s * a.dir[gi];
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
      double acc = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) acc += B[j + 6 * i] * bz[i];
      P[j + 6 * ax] = acc;
    }
  }
  bool bad = false;
  const double mv = 1e-9 * (a.vel_limit + 1.0), ma = 1e-9 * (a.acc_limit + 1.0);
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const double vx = 5 * (P[j + 1] - P[j]), vy = 5 * (P[j + 7] - P[j + 6]), vz = 5 * (P[j + 13] - P[j + 12]);
    const double d = a.vel_limit - sqrt(vx * vx + vy * vy + vz * vz) / (w * t);
    bad |= d < -mv;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double ax = 20 * (P[j + 2] - 2 * P[j + 1] + P[j]), ay = 20 * (P[j + 8] - 2 * P[j + 7] + P[j + 6]),
                 az = 20 * (P[j + 14] - 2 * P[j + 13] + P[j + 12]);
    const double d = a.acc_limit - sqrt(ax * ax + ay * ay + az * az) / (w * w * t * t);
    bad |= d < -ma;
  }
  // the 32 rungs of this row: one word, merged into the robot's mask
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  if (rung == 0 && m) atomicOr(a.mask + u, m);
}

// lay out the first batch of trial points: k=0 the current point, k=1.. the ladder step, step*0.8, step*0.8*0.8, ...
__global__ void k_ls_init(int rb, int re, const int* kmax, const double* steps_tab, const double* selfstep, int use_self,
                          const double* ptime, const double* tdir, double* step, double* tstep, double* ttime, int* done,
                          DevCounts* dc, int ls_rounds, int* bad, const unsigned* bmask) {
  int u = rb + blockIdx.x * blockDim.x + threadIdx.x;
  if (u == rb) {
    for (int r = 0; r < TOB_LS_MAXROUNDS + 1; r++) dc->ls_pending[r] = 0;
    dc->ls_rounds = ls_rounds;
  }
  if (u >= re) return;
  double s = ls_base_step(u, kmax, steps_tab, selfstep, use_self, ptime, tdir);
  int n_skip = 0;
  if (bmask) {
    // leading rungs at which a velocity / acceleration bound is violated for certain (k_ls_bound_mask): the reference's
    // energy is +inf there and its loop moves on -- so does this one, without evaluating them.  The steps are the same
    // products s * 0.8 * 0.8 ... the evaluated ladder would have formed.
    const unsigned m = bmask[u];
    const int lead = m == 0xffffffffu ? 32 : __ffs(~m) - 1;
    for (; n_skip < lead; n_skip++) s *= 0.8;
  }
  if (n_skip) atomicAdd(&dc->ls_skipped, (unsigned long long)n_skip);
  step[u] = s;
  tstep[u * TOB_LS_TRIALS] = 0.0;
  ttime[u * TOB_LS_TRIALS] = ptime[u];
  for (int k = 1; k < TOB_LS_TRIALS; k++) {
    tstep[u * TOB_LS_TRIALS + k] = s;
    ttime[u * TOB_LS_TRIALS + k] = ptime[u] + s * tdir[u];
    s *= 0.8;
  }
  done[u] = 0;
  for (int k = 0; k < TOB_LS_TRIALS; k++) bad[u * TOB_LS_TRIALS + k] = 0;   // infeasibility flags of the trial slots (barrier.cu)
}

// coupled mode (Optimization3D_multi::update_spline :585-636): ONE step and ONE piece time for all robots.
// step = min(couple_self_step, min_u position_step_u), clamped so that the shared piece time stays positive.
__global__ void k_ls_init_coupled(int U, const int* kmax, const double* steps_tab, const double* selfstep, const double* ptime,
                                  const double* tdir, double* step, double* tstep, double* ttime, int* done, DevCounts* dc,
                                  int ls_rounds) {
  if (threadIdx.x || blockIdx.x) return;
  for (int r = 0; r < TOB_LS_MAXROUNDS + 1; r++) dc->ls_pending[r] = 0;
  dc->ls_rounds = ls_rounds;
  int km = 0;
  for (int u = 0; u < U; u++) if (kmax[u] > km) km = kmax[u];
  double s = selfstep[0];
  const double sp = steps_tab[km];
  if (sp < s) s = sp;
  if (ptime[0] + s * tdir[0] <= 0) s = -0.95 * ptime[0] / tdir[0];
  for (int u = 0; u < U; u++) {
    double sk = s;
    step[u] = s;
    tstep[u * TOB_LS_TRIALS] = 0.0;
    ttime[u * TOB_LS_TRIALS] = ptime[0];
    for (int k = 1; k < TOB_LS_TRIALS; k++) {
      tstep[u * TOB_LS_TRIALS + k] = sk;
      ttime[u * TOB_LS_TRIALS + k] = ptime[0] + sk * tdir[0];
      sk *= 0.8;
    }
    done[u] = 0;
  }
}

// joint Armijo test on the summed energy (robots in index order like the reference's loop, Optimization3D_multi.h:641-657)
__global__ void k_armijo_coupled(int U, const double* etr, const double* wolfe, const double* ptime, const double* tdir, double* step,
                                 double* ptrial, double* tstep, double* ttime, int* done, int* n_active) {   // n_active: &dc->ls_pending[slot]
  if (threadIdx.x || blockIdx.x) return;
  if (done[0]) return;
  double e0 = 0;
  for (int u = 0; u < U; u++) e0 += etr[u * TOB_LS_TRIALS];
  const double w = wolfe[0];
  for (int k = 1; k < TOB_LS_TRIALS; k++) {
    const double s = tstep[k];
    double e1 = 0;
    for (int u = 0; u < U; u++) e1 += etr[u * TOB_LS_TRIALS + k];
    if (!(e0 - 1e-4 * w * s < e1)) {
      for (int u = 0; u < U; u++) { step[u] = s; ptrial[u] = ttime[k]; done[u] = 1; }
      return;
    }
  }
  double s = tstep[TOB_LS_TRIALS - 1] * 0.8;
  for (int k = 1; k < TOB_LS_TRIALS; k++) {
    for (int u = 0; u < U; u++) { tstep[u * TOB_LS_TRIALS + k] = s; ttime[u * TOB_LS_TRIALS + k] = ptime[0] + s * tdir[0]; }
    s *= 0.8;
  }
  atomicAdd(n_active, 1);
}

// robots [rb,re): piece time <- accepted trial time; the control points move only for the robots this context owns
// ([ob,oe): coupled sharded runs keep the ONE shared piece time current in every robot slot)
__global__ void k_apply_step(int rb, int re, int ob, int oe, int T, const double* step, const double* dir, const double* ptrial,
                             double* spline, double* ptime, const DevCounts* guard) {
  int u = rb + blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= re || iteration_blocked(guard)) return;
  if (i < 3 * T && u >= ob && u < oe) {
    size_t g = (size_t)u * 3 * T + i;
    spline[g] = spline[g] + step[u] * dir[g];
  }
  if (i == 0) ptime[u] = ptrial[u];
}

// sharded runs: every rank publishes its overflow / error bits, and adopts the others' so that all ranks take the same
// decision about committing the iteration (a rank that repeated the iteration alone would leave the collectives unmatched)
__global__ void k_flags_pack(const DevCounts* dc, double* all, int rank) { all[rank] = (double)dc->overflow; }
__global__ void k_flags_merge(DevCounts* dc, const double* all, int rank, int world) {
  uint32_t remote = 0;
  for (int r = 0; r < world; r++) if (r != rank) remote |= (uint32_t)all[r];
  uint32_t add = 0;
  if (remote & TOB_ERR_SOLVE) add |= TOB_ERR_SOLVE;
  if (remote & TOB_OVF_RETRY) add |= TOB_OVF_REMOTE;
  if (add) dc->overflow |= add;
}

__global__ void k_zero_steps(double* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

// pinned read-back area: at least `bytes`
static int pinned_ensure(tob_ctx* c, size_t bytes) {
  if (bytes <= c->h_pinned_bytes) return 0;
  size_t want = bytes < 65536 ? 65536 : bytes + bytes / 4;
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  c->h_pinned = nullptr; c->h_pinned_bytes = 0;
  TOB_CUDA(c, cudaMallocHost((void**)&c->h_pinned, want));
  c->h_pinned_bytes = want;
  return 0;
}

static int read_back(tob_ctx* c, const void* dev, size_t bytes) {
  TOB_TRY(pinned_ensure(c, bytes));
  TOB_CUDA(c, cudaMemcpyAsync(c->h_pinned, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// device counts -> pinned mirror; the ONE host synchronisation of an iteration
int sync_counts(tob_ctx* c) {
  TOB_CUDA(c, cudaMemcpyAsync(c->h_dc, c->dc.p, sizeof(DevCounts), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_cand = c->h_dc->n_cand;
  c->n_planes = c->h_dc->n_planes;
  return 0;
}

static int clear_overflow(tob_ctx* c) {
  TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->overflow, 0, sizeof(uint32_t), c->stream));
  return 0;
}

int grow_cand_capacity(tob_ctx* c, uint64_t need) {
  uint64_t cap = c->cand_cap ? c->cand_cap : (1u << 20);
  while (cap < need + need / 4) cap *= 2;
  if (cap > 0xfff00000ull) return fail_msg(c, "candidate count exceeds the 32-bit index range");
  c->cand_cap = cap;
  return ensure_query_buffers(c);
}

// run `launch` (asynchronous query kernels), read the counts, grow the candidate buffers and repeat on overflow
template <class F>
static int run_checked(tob_ctx* c, F&& launch) {
  for (int attempt = 0; attempt < 8; attempt++) {
    TOB_TRY(launch());
    TOB_TRY(sync_counts(c));
    if (!(c->h_dc->overflow & (TOB_OVF_CAND | TOB_OVF_LIVE))) return 0;
    const uint32_t ovf = c->h_dc->overflow;
    TOB_TRY(clear_overflow(c));
    if (ovf & TOB_OVF_CAND) TOB_TRY(grow_cand_capacity(c, c->h_dc->n_cand));
    if (ovf & TOB_OVF_LIVE) TOB_TRY(ensure_live_buffers(c, (uint64_t)c->h_dc->n_live + c->h_dc->n_new + 1));
  }
  return fail_msg(c, "candidate buffers keep overflowing");
}

// planes of robots [rb,re) from the resident splines (asynchronous)
static int separate_resident(tob_ctx* c, int rb, int re, int with_self) {
  const int U = c->n_robots();
  bool ws = with_self && U > 1;
  // inter-robot planes need every robot's rows; otherwise only the owned ones
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, ws ? 0 : rb, ws ? U : re, 1));
  TOB_TRY(broadphase(c, rb, re, c->prm.offset + c->prm.margin, 1));
  TOB_TRY(narrowphase_planes(c, rb, re, ws ? 1 : 0));
  return 0;
}

// ---- line search of robots [rb,re): s_step / trial tables / s_done prepared by k_ls_init -------------------------------
// round r >= 1 evaluates trials 1..8 (round 0 also trial 0 = the current point, the "e" of the reference)
static int ls_round(tob_ctx* c, int rb, int re, int wolfe_idx, bool coupled, int round, int slot) {
  cudaStream_t st = c->stream;
  if (!coupled) return line_search_round(c, rb, re, wolfe_idx, round == 0 ? 0 : 1, round == 0 ? c->ls_kte0 : (round == 1 ? c->ls_kte1 : c->ls_kte), slot,
                                         round == 0 && !c->ls_e0_ready ? 0 : 1);
  TOB_TRY(energy_trials(c, rb, re, c->s_dir.p, c->s_tstep.p, c->s_ttime.p, TOB_LS_TRIALS, round == 0 ? 0 : 1, TOB_LS_TRIALS, c->s_etr.p));
  TOB_TRY(exchange_robots(c, c->s_etr.p, TOB_LS_TRIALS, sizeof(double)));   // joint Armijo: every rank sums all robots in robot order
  k_armijo_coupled<<<1, 32, 0, st>>>(c->n_robots(), c->s_etr.p, c->s_wolfe.p, c->s_ptime.p, c->s_tdir.p, c->s_step.p, c->s_ptrial.p,
                                     c->s_tstep.p, c->s_ttime.p, c->s_done.p, &c->dc.p->ls_pending[slot]);
  TOB_LAUNCH_CHECK(c);
  c->ctr.line_search_trials += (uint64_t)(re - rb) * (TOB_LS_TRIALS - 1);
  return 0;
}

static int apply_step(tob_ctx* c, int rb, int re, bool guarded, bool coupled = false) {
  const int b = coupled ? 0 : rb, e = coupled ? c->n_robots() : re;
  dim3 grid(div_up(3 * c->T, 128), e - b);
  k_apply_step<<<grid, 128, 0, c->stream>>>(b, e, rb, re, c->T, c->s_step.p, c->s_dir.p, c->s_ptrial.p, c->s_spline.p, c->s_ptime.p,
                                            guarded ? c->dc.p : nullptr);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// the rounds launched ahead of the host (TOB_LS_ROUNDS x 8 ladder rungs)
static int ls_launch_ahead(tob_ctx* c, int rb, int re, int wolfe_idx, bool coupled) {
  for (int r = 0; r < c->ls_rounds; r++) TOB_TRY(ls_round(c, rb, re, wolfe_idx, coupled, r, r));
  return 0;
}

// host-driven continuation for the rare robot that needs more than TOB_LS_ROUNDS x 8 rungs; h_dc must be current
static int ls_finish(tob_ctx* c, int rb, int re, int wolfe_idx, bool coupled) {
  const int slot = TOB_LS_MAXROUNDS;   // scratch slot, re-zeroed before every extra round
  int pending = c->h_dc->ls_pending[c->ls_rounds - 1];
  for (int round = c->ls_rounds; pending > 0 && round < 400; round++) {
    TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->ls_pending[slot], 0, sizeof(int), c->stream));
    TOB_TRY(ls_round(c, rb, re, wolfe_idx, coupled, round, slot));
    TOB_TRY(sync_counts(c));
    pending = c->h_dc->ls_pending[slot];
  }
  if (pending > 0) return fail_msg(c, "line search: no Armijo rung accepted after 400 rounds of the 0.8 ladder (step < 1e-300)");
  TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->ls_pending[c->ls_rounds - 1], 0, sizeof(int), c->stream));
  return 0;
}

// complete line search with the step applied (function-level entry points)
static int line_search(tob_ctx* c, int rb, int re, int wolfe_idx, bool coupled = false) {
  TOB_TRY(ls_launch_ahead(c, rb, re, wolfe_idx, coupled));
  TOB_TRY(sync_counts(c));
  TOB_TRY(ls_finish(c, rb, re, wolfe_idx, coupled));
  return apply_step(c, rb, re, false, coupled);
}

// buffers of the whole iteration, sized from the parameters and the candidate capacity: nothing is allocated (and no size
// is read back) between the first and the last kernel of an iteration
static int ensure_iter_buffers(tob_ctx* c) {
  const size_t rows = (size_t)c->rows_all(), U = (size_t)c->n_robots(), P = (size_t)c->prm.piece_num;
  TOB_TRY(ensure_query_buffers(c));
  if (c->live_planes()) TOB_TRY(ensure_live_buffers(c, c->live_cap ? c->live_cap : 1));
  TOB_CUDA(c, c->geo.P.ensure(18 * rows)); TOB_CUDA(c, c->geo.D.ensure(18 * rows)); TOB_CUDA(c, c->geo.box.ensure(6 * rows));
  TOB_CUDA(c, c->geo.klo.ensure(TOB_KDOP_AXES * rows)); TOB_CUDA(c, c->geo.khi.ensure(TOB_KDOP_AXES * rows));
  TOB_CUDA(c, c->geo.kf.ensure(TOB_KF_ROW * rows)); TOB_CUDA(c, c->geo.kc.ensure(4 * rows));
  TOB_CUDA(c, c->row_e.ensure(TOB_EN_REC * rows * TOB_LS_TRIALS)); TOB_CUDA(c, c->row_bad.ensure(U * TOB_LS_TRIALS + 1));
  TOB_CUDA(c, c->pc_g.ensure(19 * U * P)); TOB_CUDA(c, c->pc_h.ensure(361 * U * P)); TOB_CUDA(c, c->pc_flag.ensure(U * P));
  if (U > 1 && c->cloud_n1.empty()) {   // inter-robot scratch (never used by independent problems)
    const size_t n = (size_t)c->n_tr * (U * (U - 1) / 2);
    TOB_CUDA(c, c->self_pl.ensure(4 * n + 4)); TOB_CUDA(c, c->self_ok.ensure(n + 1));
    TOB_CUDA(c, c->self_hits.ensure(2 * (n ? n : 1) + 2));
  }
  return 0;
}

// One ADMM iteration, launched without any host read-back: the state-changing tail (apply step, slack / dual update) is
// guarded on the device by dc->overflow / dc->ls_pending and the caller inspects dc afterwards.  Sharded multi-GPU runs
// take the same path: the exchanges are enqueued on the stream like kernels, and the overflow / error bits of every rank
// travel with the second exchange so that all ranks agree on whether the iteration commits.
// few rows: cover 8 rungs per launch (latency); many rows: one rung per round (throughput), more rounds ahead
static void ls_policy(tob_ctx* c, int rb, int re, bool coupled) {
  // The ladder starts behind the rungs that violate a bound for certain (k_ls_bound_mask), and then the first evaluated rung
  // is accepted by nearly every search (tob_counters.ls_rung_hist: 127.9 of 128 per iteration of a batch shard, 1.0 of 1 on the
  // forest scene): round 0 evaluates that rung alone (the energy of the current point comes from the gradient pass).
  // Few rows (latency regime): a search that goes on gets 8 rungs in the one further round launched ahead.  Many rows
  // (throughput regime, >= 8192): 2 rungs in round 1, 4 in round 2 = 7 rungs before the host has to continue a search; every
  // round launched ahead costs 14 us when nobody is left in it.  Measured (ms per iteration, kte0,kte,rounds[,kte1]):
  // batch shard 2,5,5 2.176 / 2,5,3,3 2.160 / 2,5,2,3 2.148 / 2,5,4,3 2.168; forest 9,9,2 0.324 / 5,9,2 0.315 / 2,9,2 0.304;
  // 64 UAVs 9,9,2 0.389 / 2,9,2 0.350.  Before the mask: (3,9,3) 4.89 ms of energy kernels per iteration of the whole batch,
  // (3,5,4) 4.09, (2,5,4) 3.76, (2,9,3) 5.11.
  const bool many = !coupled && (long long)(re - rb) * c->n_tr >= 8192;
  c->ls_kte0 = coupled ? TOB_LS_TRIALS : 2;
  c->ls_kte = many ? 5 : TOB_LS_TRIALS;
  c->ls_rounds = many ? 3 : 2;
  c->ls_kte1 = many ? 3 : TOB_LS_TRIALS;
  if (const char* e = getenv("TRAJOPT_B200_LS")) {          // "kte0,kte,rounds[,kte1]": tuning / experiments
    int k0 = 0, k = 0, r = 0, k1 = 0;
    const int got = coupled ? 0 : sscanf(e, "%d,%d,%d,%d", &k0, &k, &r, &k1);
    if (got >= 3 && k0 >= 2 && k0 <= TOB_LS_TRIALS && k >= 2 && k <= TOB_LS_TRIALS && r >= 1 && r <= TOB_LS_MAXROUNDS) {
      c->ls_kte0 = k0; c->ls_kte = k; c->ls_rounds = r; c->ls_kte1 = k;
      if (got == 4 && k1 >= 2 && k1 <= TOB_LS_TRIALS) c->ls_kte1 = k1;
    }
  }
}

static int iterate_launch(tob_ctx* c, int mode) {
  const int rb = c->own_begin, re = c->own_end;
  // mode 2: the robot slots hold INDEPENDENT single-UAV problems (no inter-robot terms, no exchange): everything below that
  // is conditional on "several robots" sees one robot
  const int U = mode == 2 ? 1 : c->n_robots();
  cudaStream_t st = c->stream;
  const bool coupled = mode == 1;
  const bool shard = c->sharded() && U > 1;
  // (1) planes.  The control points of the other robots are only needed for the inter-robot planes: the all-gather runs on
  // the side stream next to the obstacle pass of the owned robots (rows, broadphase, 49-DOP + GJK)
  if (shard) {
    TOB_TRY(exchange_fork(c));
    TOB_TRY(exchange_robots(c, c->s_spline.p, (size_t)3 * c->T, sizeof(double)));
    TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, rb, re, 1));
    TOB_TRY(broadphase(c, rb, re, c->prm.offset + c->prm.margin, 1));
    TOB_TRY(narrowphase_planes(c, rb, re, -1));
    TOB_TRY(exchange_join(c));
    if (rb > 0) TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, 0, rb, 1));
    if (re < U) TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, re, U, 1));
    TOB_TRY(narrowphase_finish(c, rb, re, 1));
  } else {
    TOB_TRY(separate_resident(c, rb, re, U > 1));
  }
  // (2) Newton direction (geo.P of the owned rows is still current from the plane pass)
  TOB_TRY(gradient_blocks(c, rb, re, 1));
  if (coupled) TOB_TRY(solve_coupled(c));
  else TOB_TRY(solve_directions(c, rb, re, U > 1));
  // (3) CCD step bound
  if (U > 1) {
    if (shard) {
      // one grouped exchange on the side stream: directions (+ wolfe, gnorm of the decoupled solves) and every rank's
      // overflow / error bits; next to it the obstacle CCD of the owned robots, which needs their own directions only
      k_flags_pack<<<1, 1, 0, st>>>(c->dc.p, c->ovf_all.p, c->comm_rank);
      TOB_LAUNCH_CHECK(c);
      TOB_TRY(exchange_fork(c));
      TOB_TRY(exchange_group_begin(c));
      TOB_TRY(exchange_robots(c, c->s_dir.p, (size_t)3 * c->T, sizeof(double)));
      if (!coupled) {
        TOB_TRY(exchange_robots(c, c->s_wolfe.p, 1, sizeof(double)));
        TOB_TRY(exchange_robots(c, c->s_gnorm.p, 1, sizeof(double)));
      }
      TOB_TRY(exchange_ranks(c, c->ovf_all.p));
      TOB_TRY(exchange_group_end(c));
      TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, rb, re, 3));
      TOB_TRY(ccd_position_steps(c, rb, re));
      TOB_TRY(exchange_join(c));
      k_flags_merge<<<1, 1, 0, st>>>(c->dc.p, c->ovf_all.p, c->comm_rank, c->comm_world);
      TOB_LAUNCH_CHECK(c);
      if (rb > 0) TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, 0, rb, 3));
      if (re < U) TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, re, U, 3));
    } else {
      TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, 0, U, 3));
    }
    TOB_TRY(self_ccd_steps(c, coupled ? 1 : 0, c->s_selfstep.p));
    if (!shard) TOB_TRY(ccd_position_steps(c, rb, re));
  } else {
    TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, rb, re, 3));
    TOB_TRY(ccd_position_steps(c, rb, re));
  }
  // (4) Armijo; multi-robot: every robot uses the LAST robot's wolfe (global overwritten, Optimization3D_multi.h:730,792)
  int wolfe_idx = -1;
  ls_policy(c, rb, re, coupled);
  if (coupled) {
    if (U == 1) { TOB_CUDA(c, cudaMemcpyAsync(c->s_selfstep.p, c->d_steps.p, sizeof(double), cudaMemcpyDeviceToDevice, st)); }   // 0.8^0 = 1
    if (shard) TOB_TRY(exchange_robots(c, c->kmax.p, 1, sizeof(int)));   // shared step = min over ALL robots' position steps
    k_ls_init_coupled<<<1, 32, 0, st>>>(c->n_robots(), c->kmax.p, c->d_steps.p, c->s_selfstep.p, c->s_ptime.p, c->s_tdir.p, c->s_step.p,
                                        c->s_tstep.p, c->s_ttime.p, c->s_done.p, c->dc.p, c->ls_rounds);
    TOB_LAUNCH_CHECK(c);
    wolfe_idx = 0;
  } else {
    // largest step at which every plane still has all its control points on the right side (one pass over the planes)
    const unsigned* bmask = nullptr;
    {
      const char* e = getenv("TRAJOPT_B200_LS_SKIP");      // 0: every rung of the ladder is evaluated (tests, A/B)
      if (!e || atoi(e) != 0) {
        TOB_CUDA(c, cudaMemsetAsync(c->s_lsmask.p + rb, 0, (size_t)(re - rb) * sizeof(unsigned), st));
        LsMaskArgs m;
        m.rb = rb; m.re = re; m.n_tr = c->n_tr; m.res = c->prm.res; m.T = c->T; m.use_self = U > 1 ? 1 : 0;
        m.kmax = c->kmax.p; m.steps_tab = c->d_steps.p; m.selfstep = c->s_selfstep.p; m.ptime = c->s_ptime.p; m.tdir = c->s_tdir.p;
        m.spline = c->s_spline.p; m.dir = c->s_dir.p; m.basis = c->d_basis.p; m.weight = c->d_weight.p;
        m.vel_limit = c->prm.vel_limit; m.acc_limit = c->prm.acc_limit; m.mask = c->s_lsmask.p;
        k_ls_bound_mask<<<div_up((re - rb) * c->n_tr, 4), 128, 0, st>>>(m);
        TOB_LAUNCH_CHECK(c);
        bmask = c->s_lsmask.p;
      }
    }
    k_ls_init<<<div_up(re - rb, 64), 64, 0, st>>>(rb, re, c->kmax.p, c->d_steps.p, c->s_selfstep.p, U > 1 ? 1 : 0, c->s_ptime.p,
                                                  c->s_tdir.p, c->s_step.p, c->s_tstep.p, c->s_ttime.p, c->s_done.p, c->dc.p,
                                                  c->ls_rounds, c->row_bad.p, bmask);
    TOB_LAUNCH_CHECK(c);
    wolfe_idx = U > 1 ? U - 1 : -1;
  }
  c->ls_e0_ready = !coupled;         // gradient_blocks ran at this very point with these planes: slot 0 holds E(x)
  TOB_TRY(ls_launch_ahead(c, rb, re, wolfe_idx, coupled));
  // (5) step, slack + dual: guarded on the device (see iterate_once)
  TOB_TRY(apply_step(c, rb, re, true, coupled));
  TOB_TRY(slack_update(c, rb, re, 1));
  return 0;
}

static void graph_drop(tob_ctx* c) {
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr;
  c->graph_mode = -1;
}

// Launch one deferred iteration: through the captured CUDA graph when the buffers have not moved since the capture.
static int iterate_submit(tob_ctx* c, int mode) {
  // NCCL calls are captured like kernels; the legacy callback exchange (a host function enqueuing work through another
  // library) and the per-launch profiling events are not
  const bool can_graph = c->use_graph && !c->prof_on && !c->ag && (mode == 0 || mode == 2 || (mode == 1 && c->n_robots() > 1));
  if (!can_graph) return iterate_launch(c, mode);
  if (c->graph_exec && (c->graph_gen != alloc_generation() || c->graph_mode != mode)) graph_drop(c);
  if (!c->graph_exec) {
    // capture only a launch sequence that a plain run has already executed without allocating
    const unsigned long long gen0 = alloc_generation();
    TOB_TRY(ensure_iter_buffers(c));
    if (gen0 != alloc_generation() || c->graph_warm_gen != gen0) {
      c->graph_warm_gen = alloc_generation();
      TOB_TRY(iterate_launch(c, mode));
      if (c->graph_warm_gen != alloc_generation()) c->graph_warm_gen = ~0ull;   // allocated on the way: warm up again
      return 0;
    }
    cudaGraph_t g = nullptr;
    const uint64_t l0 = c->ctr.kernel_launches;
    TOB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    int rc = iterate_launch(c, mode);
    c->capturing = false;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->graph_nodes = c->ctr.kernel_launches - l0;
    c->ctr.kernel_launches = l0;
    if (rc || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      c->use_graph = false;                      // fall back to plain stream launches for good
      return iterate_launch(c, mode);
    }
    e = cudaGraphInstantiate(&c->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { c->graph_exec = nullptr; c->use_graph = false; cudaGetLastError(); return iterate_launch(c, mode); }
    c->graph_gen = alloc_generation();
    c->graph_mode = mode;
  }
  TOB_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
  c->ctr.kernel_launches += c->graph_nodes;
  c->ctr.line_search_trials += (uint64_t)(c->own_end - c->own_begin) * ((c->ls_kte0 - 1) + (c->ls_rounds > 1 ? c->ls_kte1 - 1 : 0) + (c->ls_kte - 1) * (c->ls_rounds > 2 ? c->ls_rounds - 2 : 0));
  if (c->n_robots() > 1 && mode != 2) c->ctr.self_pairs += (uint64_t)2 * c->n_tr * (c->n_robots() * (c->n_robots() - 1) / 2);
  return 0;
}

static int iterate_once(tob_ctx* c, int mode, double* gnorm_out) {
  const int U = c->n_robots(), rb = c->own_begin, re = c->own_end;
  const bool coupled = mode == 1;
  if (mode < 0 || mode > 2) return fail_msg(c, "tob_admm_iterate: mode must be 0 (decoupled), 1 (coupled) or 2 (independent problems)");
  if (mode != 2 && !c->cloud_n1.empty() && U > 1) return fail_msg(c, "per-robot clouds (tob_cloud_upload_batch) go with mode 2 (independent problems)");
  if (coupled && c->ag) return fail_msg(c, "coupled mode over several GPUs needs the native NCCL exchange (tob_nccl_init_rank / tob_nccl_attach), not the tob_set_shard callbacks");
  if (mode == 2 && c->sharded() && (rb != 0 || re != U)) return fail_msg(c, "independent problems (mode 2) are not sharded: give every GPU its own context and problems");
  TOB_TRY(ensure_iter_buffers(c));
  const int wolfe_idx = coupled ? 0 : ((U > 1 && mode != 2) ? U - 1 : -1);
  for (int attempt = 0;; attempt++) {
    const uint32_t done0 = c->h_dc->iters_done;
    TOB_TRY(iterate_submit(c, mode));
    TOB_CUDA(c, cudaMemcpyAsync(c->h_pinned, c->s_gnorm.p, U * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TOB_TRY(sync_counts(c));
    if (c->h_dc->iters_done != done0) break;                 // committed on the device
    const uint32_t ovf = c->h_dc->overflow;
    if (ovf & TOB_OVF_RETRY) {                               // nothing was changed: grow and run the iteration again
      if (attempt >= 8) return fail_msg(c, "candidate buffers keep overflowing");
      TOB_TRY(clear_overflow(c));
      if (ovf & TOB_OVF_CAND) TOB_TRY(grow_cand_capacity(c, c->h_dc->n_cand));
      if (ovf & TOB_OVF_LIVE) TOB_TRY(ensure_live_buffers(c, (uint64_t)c->h_dc->n_live + c->h_dc->n_new + 1));
      continue;
    }
    if (ovf & TOB_ERR_SOLVE) {                               // same on every rank of a sharded run
      TOB_TRY(clear_overflow(c));
      return fail_msg(c, "Newton matrix is not positive definite (Cholesky pivot or Schur complement <= 0): iteration not committed");
    }
    // a robot needs more than the rungs launched ahead: finish its search from the host, then commit.  (Decoupled: the
    // search is rank-local.  Coupled: every rank holds the same energies and takes the same decisions, exchanges matched.)
    TOB_TRY(ls_finish(c, rb, re, wolfe_idx, coupled));
    TOB_TRY(apply_step(c, rb, re, false, coupled));
    TOB_TRY(slack_update(c, rb, re, 0));
    break;
  }
  // gnorm global of the reference
  if (gnorm_out) {
    double g = 0;
    const double* hp = c->h_pinned;
    if (U == 1) g = hp[0];
    else { for (int u = 0; u < U; u++) g += hp[u]; g /= double(U); }   // mode 2: mean over the independent problems
    *gnorm_out = g;
  }
  return 0;
}

}  // namespace tob

using namespace tob;

extern "C" {

int tob_ctx_create(int device, tob_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail_msg(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail_msg(nullptr, "bad device index");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, "cudaSetDevice", e, __FILE__, __LINE__);
  tob_ctx* c = new tob_ctx();
  c->device = device;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  c->sm_count = prop.multiProcessorCount; c->cc_major = prop.major; c->cc_minor = prop.minor;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return fail(nullptr, "cudaStreamCreate", e, __FILE__, __LINE__); }
  if ((e = cudaMallocHost((void**)&c->h_pinned, 65536)) != cudaSuccess) { delete c; return fail(nullptr, "cudaMallocHost", e, __FILE__, __LINE__); }
  c->h_pinned_bytes = 65536;
  if ((e = c->red.ensure(4096)) != cudaSuccess) { delete c; return fail(nullptr, "cudaMalloc", e, __FILE__, __LINE__); }
  if ((e = cudaMallocHost((void**)&c->h_dc, sizeof(DevCounts))) != cudaSuccess) { delete c; return fail(nullptr, "cudaMallocHost", e, __FILE__, __LINE__); }
  memset(c->h_dc, 0, sizeof(DevCounts));
  if ((e = cudaMalloc((void**)&c->dc.p, sizeof(DevCounts))) != cudaSuccess) { delete c; return fail(nullptr, "cudaMalloc", e, __FILE__, __LINE__); }
  c->dc.cap = 1;
  cudaMemset(c->dc.p, 0, sizeof(DevCounts));
  if (const char* g = getenv("TRAJOPT_B200_NO_GRAPH")) c->use_graph = !(g[0] && g[0] != '0');
  std::vector<double> steps(TOB_LADDER + 2);
  double s = 1.0;
  for (int k = 0; k < TOB_LADDER + 2; k++) { steps[k] = s; s *= 0.8; }
  c->d_steps.ensure(steps.size());
  cudaMemcpy(c->d_steps.p, steps.data(), steps.size() * sizeof(double), cudaMemcpyHostToDevice);
  *out = c;
  return 0;
}

void tob_ctx_destroy(tob_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  delete c;   // ~tob_ctx: graph, events, pinned areas, stream; every DBuf frees its own allocation
}

const char* tob_last_error(const tob_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int tob_device_info(const tob_ctx* c, int* sm_count, int* cc_major, int* cc_minor) {
  if (!c) return 1;
  if (sm_count) *sm_count = c->sm_count;
  if (cc_major) *cc_major = c->cc_major;
  if (cc_minor) *cc_minor = c->cc_minor;
  return 0;
}

int tob_set_params(tob_ctx* c, const tob_params* p) {
  if (!c || !p) return 1;
  if (p->piece_num < 1 || p->res < 1 || p->res > 16 || p->uav_num < 1) return fail_msg(c, "tob_set_params: bad sizes");
  cudaSetDevice(c->device);
  c->prm = *p;
  c->prm.optimal_plane = p->optimal_plane ? 1 : 0;
  c->n_tr = p->piece_num * p->res;
  c->T = 6 + 3 * (p->piece_num - 1);
  c->have_params = true;
  c->have_tables = false;
  c->states_valid = false;
  c->own_begin = 0; c->own_end = p->uav_num;
  if (c->nccl_comm) {
    if (c->comm_world > p->uav_num) return fail_msg(c, "tob_set_params: fewer robots than ranks of the attached communicator");
    shard_partition(c);
  } else { c->ag = nullptr; c->ar = nullptr; c->comm_rank = 0; c->comm_world = 1; }
  c->n_planes = 0;
  if (!c->cloud_n1.empty()) {      // per-robot clouds are tied to the row layout: upload them again
    c->cloud_n1.clear(); c->cloud_l1.clear(); c->h_row_task.clear();
    c->n_pts = 0;
  }
  TOB_TRY(reset_live_planes(c));     // is_seperate / is_self_seperate start empty (init_variable)
  return alloc_states(c);
}

static int upload_tables(tob_ctx* c) {
  TOB_CUDA(c, c->d_basis.ensure(c->h_basis.size())); TOB_CUDA(c, c->d_weight.ensure(c->h_weight.size()));
  TOB_CUDA(c, c->d_convert.ensure(c->h_convert.size())); TOB_CUDA(c, c->d_mdyn.ensure(36)); TOB_CUDA(c, c->d_kdop.ensure(147));
  TOB_TRY(upload(c, c->d_basis, c->h_basis.data(), c->h_basis.size()));
  TOB_TRY(upload(c, c->d_weight, c->h_weight.data(), c->h_weight.size()));
  TOB_TRY(upload(c, c->d_convert, c->h_convert.data(), c->h_convert.size()));
  TOB_TRY(upload(c, c->d_mdyn, c->h_mdyn.data(), 36));
  TOB_TRY(upload(c, c->d_kdop, c->h_kdop.data(), 147));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->have_tables = true;
  return 0;
}

int tob_make_tables(tob_ctx* c, const double* time_weight) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  make_tables_host(c->prm, time_weight, c->h_basis, c->h_weight, c->h_convert, c->h_mdyn, c->h_kdop);
  return upload_tables(c);
}

int tob_set_tables(tob_ctx* c, const double* basis, const double* weight, const double* convert, const double* mdyn,
                   const double* kdop) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  c->h_basis.assign(basis, basis + (size_t)36 * c->n_tr);
  c->h_weight.assign(weight, weight + c->n_tr);
  c->h_convert.assign(convert, convert + (size_t)36 * c->prm.piece_num);
  c->h_mdyn.assign(mdyn, mdyn + 36);
  c->h_kdop.assign(kdop, kdop + 147);
  return upload_tables(c);
}

int tob_get_tables(const tob_ctx* c, double* basis, double* weight, double* convert, double* mdyn, double* kdop) {
  if (!c || !c->have_tables) return 1;
  if (basis) memcpy(basis, c->h_basis.data(), c->h_basis.size() * sizeof(double));
  if (weight) memcpy(weight, c->h_weight.data(), c->h_weight.size() * sizeof(double));
  if (convert) memcpy(convert, c->h_convert.data(), c->h_convert.size() * sizeof(double));
  if (mdyn) memcpy(mdyn, c->h_mdyn.data(), 36 * sizeof(double));
  if (kdop) memcpy(kdop, c->h_kdop.data(), 147 * sizeof(double));
  return 0;
}

int tob_cloud_upload(tob_ctx* c, const double* V, uint32_t n) {
  if (!c || !V) return 1;
  cudaSetDevice(c->device);
  TOB_TRY(reset_live_planes(c));     // live planes are keyed by the position in the sorted cloud
  return lbvh_build(c, V, n);
}
int tob_cloud_upload_batch(tob_ctx* c, const double* const* V, const uint32_t* n, int n_clouds) {
  if (!c || !V || !n) return 1;
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  TOB_TRY(reset_live_planes(c));
  return lbvh_build_batch(c, V, n, n_clouds);
}
uint32_t tob_cloud_size(const tob_ctx* c) { return c ? c->n_pts : 0; }

// ---- broadphase --------------------------------------------------------------------------------------------------
static int bp_download(tob_ctx* c, int n_robots, uint32_t* offsets, uint32_t* ids, uint64_t cap, uint64_t* total) {
  const int rows = n_robots * c->n_tr;
  uint64_t nc = c->n_cand;   // host mirror: the caller went through run_checked()
  if (total) *total = nc;
  std::vector<uint32_t> off(c->rows_all() + 1);
  TOB_CUDA(c, cudaMemcpyAsync(off.data(), c->row_off.p, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  std::vector<uint32_t> pts(nc);
  if (nc) TOB_CUDA(c, cudaMemcpyAsync(pts.data(), c->cand_pt.p, nc * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r <= rows; r++) offsets[r] = off[r];
  if (nc > cap || !ids) return 0;
  for (uint64_t i = 0; i < nc; i++) ids[i] = c->h_pid[pts[i]];
  for (int r = 0; r < rows; r++) std::sort(ids + off[r], ids + off[r + 1]);
  return 0;
}

static int stage_splines(tob_ctx* c, const double* splines, const double* dirs, int n_robots) {
  if (n_robots < 1 || n_robots > c->n_robots()) return fail_msg(c, "n_robots exceeds tob_params.uav_num");
  TOB_TRY(upload(c, c->s_spline, splines, (size_t)n_robots * 3 * c->T));
  if (dirs) TOB_TRY(upload(c, c->s_dir, dirs, (size_t)n_robots * 3 * c->T));
  c->states_valid = false;
  return 0;
}

int tob_broadphase_dcd(tob_ctx* c, const double* splines, int n_robots, double d, uint32_t* offsets, uint32_t* ids,
                       uint64_t cap, uint64_t* total) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  TOB_TRY(stage_splines(c, splines, nullptr, n_robots));
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, 0, n_robots, 0));
  TOB_TRY(run_checked(c, [&]() { return broadphase(c, 0, n_robots, d, 0); }));
  return bp_download(c, n_robots, offsets, ids, cap, total);
}

int tob_broadphase_ccd(tob_ctx* c, const double* splines, const double* directions, int n_robots, double d,
                       uint32_t* offsets, uint32_t* ids, uint64_t cap, uint64_t* total) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  TOB_TRY(stage_splines(c, splines, directions, n_robots));
  TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, 0, n_robots, 2));
  TOB_TRY(run_checked(c, [&]() { return broadphase(c, 0, n_robots, d, 0); }));
  return bp_download(c, n_robots, offsets, ids, cap, total);
}

int tob_box_query(tob_ctx* c, const double* lo, const double* hi, double d, uint32_t* ids, uint64_t cap, uint64_t* total) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  TOB_CUDA(c, c->geo.box.ensure((size_t)6 * c->rows_all()));
  double b[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
  TOB_TRY(upload(c, c->geo.box, b, 6));
  c->states_valid = false;
  TOB_TRY(run_checked(c, [&]() { return broadphase_rows(c, 0, 1, d, 0); }));
  uint64_t t = c->n_cand;
  if (total) *total = t;
  if (!ids || t > cap || t == 0) return 0;
  std::vector<uint32_t> pts(t);
  TOB_CUDA(c, cudaMemcpyAsync(pts.data(), c->cand_pt.p, t * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (uint64_t i = 0; i < t; i++) ids[i] = c->h_pid[pts[i]];
  std::sort(ids, ids + t);
  return 0;
}

int tob_edge_validity_batch(tob_ctx* c, const double* edges, int n, double d, uint8_t* valid) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  if (!edges || !valid) return fail_msg(c, "tob_edge_validity_batch: null argument");
  return edge_validity(c, edges, n, d, valid);
}

}  // extern "C"

// ---- batched primitives --------------------------------------------------------------------------------------------
template <int NA, int NB>
__global__ void k_gjk_batch(const double* A, const double* B, int n, double* v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[NA][3], b[NB][3];
  for (int j = 0; j < NA; j++) for (int k = 0; k < 3; k++) a[j][k] = A[(size_t)i * NA * 3 + k * NA + j];
  for (int j = 0; j < NB; j++) for (int k = 0; k < 3; k++) b[j][k] = B[(size_t)i * NB * 3 + k * NB + j];
  gjk_witness<NA, NB>(a, b, v + 3 * (size_t)i);
}

#define TOB_DYN_MAX 16
__global__ void k_gjk_batch_n(const double* A, int na, const double* B, int nb, int n, double* v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[TOB_DYN_MAX][3], b[TOB_DYN_MAX][3];
  for (int j = 0; j < na; j++) for (int k = 0; k < 3; k++) a[j][k] = A[(size_t)i * na * 3 + k * na + j];
  for (int j = 0; j < nb; j++) for (int k = 0; k < 3; k++) b[j][k] = B[(size_t)i * nb * 3 + k * nb + j];
  gjk_witness_n(a, na, b, nb, v + 3 * (size_t)i);
}

__global__ void k_kdop_batch_n(const double* A, int na, const double* B, int nb, const double* kdop, int n, double d, uint8_t* flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[TOB_DYN_MAX][3], b[TOB_DYN_MAX][3];
  for (int j = 0; j < na; j++) for (int k = 0; k < 3; k++) a[j][k] = A[(size_t)i * na * 3 + k * na + j];
  for (int j = 0; j < nb; j++) for (int k = 0; k < 3; k++) b[j][k] = B[(size_t)i * nb * 3 + k * nb + j];
  flags[i] = kdop_overlap_n(a, na, b, nb, kdop, d) ? 1 : 0;
}

__global__ void k_kdop_batch(const double* P, const double* q, const double* kdop, int n, double d, uint8_t* flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], b[1][3];
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) a[j][k] = P[(size_t)i * 18 + k * 6 + j];
  for (int k = 0; k < 3; k++) b[0][k] = q[(size_t)i * 3 + k];
  flags[i] = kdop_overlap<6, 1>(a, b, kdop, d) ? 1 : 0;
}

__global__ void k_plane_point_batch(const double* P, const double* q, int n, double dist, double offset, uint8_t* ok, double* c,
                                    double* d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], qq[3], cc[3] = {0, 0, 0}, dd = 0;
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) a[j][k] = P[(size_t)i * 18 + k * 6 + j];
  for (int k = 0; k < 3; k++) qq[k] = q[(size_t)i * 3 + k];
  bool r = plane_point(a, qq, dist, offset, cc, &dd);
  ok[i] = r;
  c[3 * (size_t)i] = cc[0]; c[3 * (size_t)i + 1] = cc[1]; c[3 * (size_t)i + 2] = cc[2];
  d[i] = dd;
}

__global__ void k_plane_hulls_batch(const double* P0, const double* P1, int n, double dist, double offset, double margin,
                                    int refine, uint8_t* ok, double* c, double* d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], b[6][3], cc[3] = {0, 0, 0}, dd = 0;
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) { a[j][k] = P0[(size_t)i * 18 + k * 6 + j]; b[j][k] = P1[(size_t)i * 18 + k * 6 + j]; }
  bool r = plane_hulls(a, b, dist, cc, &dd);
  if (r && refine) refine_d(a, b, cc, offset, margin, &dd, 10000);
  ok[i] = r;
  c[3 * (size_t)i] = cc[0]; c[3 * (size_t)i + 1] = cc[1]; c[3 * (size_t)i + 2] = cc[2];
  d[i] = dd;
}

__global__ void k_optimal_cd_batch(const double* P, const double* q, int n, double offset, double margin, double* cc, double* d,
                                   uint8_t* capped) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], qq[3], c3[3] = {cc[3 * (size_t)i], cc[3 * (size_t)i + 1], cc[3 * (size_t)i + 2]}, dd = d[i];
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) a[j][k] = P[(size_t)i * 18 + k * 6 + j];
  for (int k = 0; k < 3; k++) qq[k] = q[(size_t)i * 3 + k];
  capped[i] = (uint8_t)optimal_cd(a, qq, offset, margin, c3, &dd);
  cc[3 * (size_t)i] = c3[0]; cc[3 * (size_t)i + 1] = c3[1]; cc[3 * (size_t)i + 2] = c3[2];
  d[i] = dd;
}

__global__ void k_self_optimal_cd_batch(const double* P0, const double* P1, int n, double offset, double margin, double* cc, double* d,
                                        uint8_t* capped) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], b[6][3], c3[3] = {cc[3 * (size_t)i], cc[3 * (size_t)i + 1], cc[3 * (size_t)i + 2]}, dd = d[i];
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) { a[j][k] = P0[(size_t)i * 18 + k * 6 + j]; b[j][k] = P1[(size_t)i * 18 + k * 6 + j]; }
  capped[i] = (uint8_t)self_optimal_cd(a, b, offset, margin, c3, &dd);
  cc[3 * (size_t)i] = c3[0]; cc[3 * (size_t)i + 1] = c3[1]; cc[3 * (size_t)i + 2] = c3[2];
  d[i] = dd;
}

__global__ void k_refine_d_batch(const double* P0, const double* P1, const double* cc, int n, double offset, double margin, double* d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[6][3], b[6][3], c3[3] = {cc[3 * (size_t)i], cc[3 * (size_t)i + 1], cc[3 * (size_t)i + 2]}, dd = d[i];
  for (int j = 0; j < 6; j++) for (int k = 0; k < 3; k++) { a[j][k] = P0[(size_t)i * 18 + k * 6 + j]; b[j][k] = P1[(size_t)i * 18 + k * 6 + j]; }
  refine_d(a, b, c3, offset, margin, &dd, 10000);
  d[i] = dd;
}

extern "C" {

int tob_gjk_batch(tob_ctx* c, const double* A, int na, const double* B, int nb, int n, double* v) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  size_t sa = (size_t)n * na * 3, sb = (size_t)n * nb * 3;
  TOB_CUDA(c, c->scratch.ensure(sa + sb)); TOB_CUDA(c, c->scratch2.ensure(3 * (size_t)n));
  TOB_TRY(upload(c, c->scratch, A, sa)); TOB_TRY(upload(c, c->scratch, B, sb, sa));
  int g = div_up(n, 64);
  const double *dA = c->scratch.p, *dB = c->scratch.p + sa;
  if (na == 6 && nb == 1) k_gjk_batch<6, 1><<<g, 64, 0, c->stream>>>(dA, dB, n, c->scratch2.p);
  else if (na == 12 && nb == 1) k_gjk_batch<12, 1><<<g, 64, 0, c->stream>>>(dA, dB, n, c->scratch2.p);
  else if (na == 6 && nb == 6) k_gjk_batch<6, 6><<<g, 64, 0, c->stream>>>(dA, dB, n, c->scratch2.p);
  else if (na == 12 && nb == 12) k_gjk_batch<12, 12><<<g, 64, 0, c->stream>>>(dA, dB, n, c->scratch2.p);
  else if (na >= 1 && nb >= 1 && na <= TOB_DYN_MAX && nb <= TOB_DYN_MAX) k_gjk_batch_n<<<g, 64, 0, c->stream>>>(dA, na, dB, nb, n, c->scratch2.p);
  else return fail_msg(c, "tob_gjk_batch: vertex counts must be in 1..16");
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(v, c->scratch2.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_kdop_dcd_batch(tob_ctx* c, const double* P, const double* q, int n, double d, uint8_t* flags) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  TOB_CUDA(c, c->scratch.ensure((size_t)21 * n)); TOB_CUDA(c, c->scratch8.ensure(n));
  TOB_TRY(upload(c, c->scratch, P, (size_t)18 * n)); TOB_TRY(upload(c, c->scratch, q, (size_t)3 * n, (size_t)18 * n));
  k_kdop_batch<<<div_up(n, 64), 64, 0, c->stream>>>(c->scratch.p, c->scratch.p + (size_t)18 * n, c->d_kdop.p, n, d, c->scratch8.p);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(flags, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_kdop_batch(tob_ctx* c, const double* A, int na, const double* B, int nb, int n, double d, uint8_t* flags) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (na < 1 || nb < 1 || na > TOB_DYN_MAX || nb > TOB_DYN_MAX) return fail_msg(c, "tob_kdop_batch: vertex counts must be in 1..16");
  size_t sa = (size_t)n * na * 3, sb = (size_t)n * nb * 3;
  TOB_CUDA(c, c->scratch.ensure(sa + sb)); TOB_CUDA(c, c->scratch8.ensure(n));
  TOB_TRY(upload(c, c->scratch, A, sa)); TOB_TRY(upload(c, c->scratch, B, sb, sa));
  k_kdop_batch_n<<<div_up(n, 64), 64, 0, c->stream>>>(c->scratch.p, na, c->scratch.p + sa, nb, c->d_kdop.p, n, d, c->scratch8.p);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(flags, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_plane_point_batch(tob_ctx* c, const double* P, const double* q, int n, double distance, uint8_t* ok, double* cc,
                          double* d) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  TOB_CUDA(c, c->scratch.ensure((size_t)21 * n)); TOB_CUDA(c, c->scratch2.ensure((size_t)4 * n)); TOB_CUDA(c, c->scratch8.ensure(n));
  TOB_TRY(upload(c, c->scratch, P, (size_t)18 * n)); TOB_TRY(upload(c, c->scratch, q, (size_t)3 * n, (size_t)18 * n));
  k_plane_point_batch<<<div_up(n, 64), 64, 0, c->stream>>>(c->scratch.p, c->scratch.p + (size_t)18 * n, n, distance, c->prm.offset,
                                                         c->scratch8.p, c->scratch2.p, c->scratch2.p + (size_t)3 * n);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(ok, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(cc, c->scratch2.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(d, c->scratch2.p + (size_t)3 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_plane_hulls_batch(tob_ctx* c, const double* P0, const double* P1, int n, double distance, int refine, uint8_t* ok,
                          double* cc, double* d) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  TOB_CUDA(c, c->scratch.ensure((size_t)36 * n)); TOB_CUDA(c, c->scratch2.ensure((size_t)4 * n)); TOB_CUDA(c, c->scratch8.ensure(n));
  TOB_TRY(upload(c, c->scratch, P0, (size_t)18 * n)); TOB_TRY(upload(c, c->scratch, P1, (size_t)18 * n, (size_t)18 * n));
  k_plane_hulls_batch<<<div_up(n, 64), 64, 0, c->stream>>>(c->scratch.p, c->scratch.p + (size_t)18 * n, n, distance, c->prm.offset,
                                                         c->prm.margin, refine, c->scratch8.p, c->scratch2.p,
                                                         c->scratch2.p + (size_t)3 * n);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(ok, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(cc, c->scratch2.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(d, c->scratch2.p + (size_t)3 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_refine_d_batch(tob_ctx* c, const double* P0, const double* P1, const double* cc, int n, double* d_io) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  TOB_CUDA(c, c->scratch.ensure((size_t)36 * n)); TOB_CUDA(c, c->scratch2.ensure((size_t)4 * n));
  TOB_TRY(upload(c, c->scratch, P0, (size_t)18 * n)); TOB_TRY(upload(c, c->scratch, P1, (size_t)18 * n, (size_t)18 * n));
  TOB_TRY(upload(c, c->scratch2, cc, (size_t)3 * n)); TOB_TRY(upload(c, c->scratch2, d_io, n, (size_t)3 * n));
  k_refine_d_batch<<<div_up(n, 64), 64, 0, c->stream>>>(c->scratch.p, c->scratch.p + (size_t)18 * n, c->scratch2.p, n, c->prm.offset,
                                                      c->prm.margin, c->scratch2.p + (size_t)3 * n);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(d_io, c->scratch2.p + (size_t)3 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- planes ------------------------------------------------------------------------------------------------------------
int tob_optimal_cd_batch(tob_ctx* c, const double* P, const double* q, int n, double* cc, double* d_io, uint8_t* capped) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  if (n <= 0) return 0;
  TOB_CUDA(c, c->scratch.ensure((size_t)n * 25 + 8));
  TOB_CUDA(c, c->scratch8.ensure(n));
  double *dP = c->scratch.p, *dq = dP + (size_t)18 * n, *dc3 = dq + (size_t)3 * n, *dd = dc3 + (size_t)3 * n;
  TOB_CUDA(c, cudaMemcpyAsync(dP, P, (size_t)18 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(dq, q, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(dc3, cc, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(dd, d_io, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_optimal_cd_batch<<<div_up(n, 64), 64, 0, c->stream>>>(dP, dq, n, c->prm.offset, c->prm.margin, dc3, dd, c->scratch8.p);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(cc, dc3, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(d_io, dd, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (capped) TOB_CUDA(c, cudaMemcpyAsync(capped, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_self_optimal_cd_batch(tob_ctx* c, const double* P0, const double* P1, int n, double* cc, double* d_io, uint8_t* capped) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  if (n <= 0) return 0;
  TOB_CUDA(c, c->scratch.ensure((size_t)n * 40 + 8));
  TOB_CUDA(c, c->scratch8.ensure(n));
  double *d0 = c->scratch.p, *d1 = d0 + (size_t)18 * n, *dc3 = d1 + (size_t)18 * n, *dd = dc3 + (size_t)3 * n;
  TOB_CUDA(c, cudaMemcpyAsync(d0, P0, (size_t)18 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(d1, P1, (size_t)18 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(dc3, cc, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(dd, d_io, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_self_optimal_cd_batch<<<div_up(n, 64), 64, 0, c->stream>>>(d0, d1, n, c->prm.offset, c->prm.margin, dc3, dd, c->scratch8.p);
  TOB_LAUNCH_CHECK(c);
  TOB_CUDA(c, cudaMemcpyAsync(cc, dc3, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(d_io, dd, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (capped) TOB_CUDA(c, cudaMemcpyAsync(capped, c->scratch8.p, n, cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_planes_reset(tob_ctx* c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  return reset_live_planes(c);
}

int tob_live_planes(tob_ctx* c, uint32_t* rows, uint32_t* ids, double* cc, double* dd, uint64_t cap, uint64_t* total) {
  TOB_TRY(need(c, false, false));
  cudaSetDevice(c->device);
  TOB_TRY(sync_counts(c));
  const uint64_t n = c->live_key.p ? c->h_dc->n_live : 0;
  if (total) *total = n;
  if (!n || n > cap || !rows || !ids || !cc || !dd) return 0;
  std::vector<unsigned long long> key(n);
  std::vector<double> pl(4 * n);
  TOB_CUDA(c, cudaMemcpy(key.data(), c->live_key.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  TOB_CUDA(c, cudaMemcpy(pl.data(), c->live_pl.p, 4 * n * sizeof(double), cudaMemcpyDeviceToHost));
  for (uint64_t k = 0; k < n; k++) {
    const uint32_t p = (uint32_t)(key[k] & 0xffffffffu);
    rows[k] = (uint32_t)(key[k] >> 32);
    ids[k] = p < c->h_pid.size() ? c->h_pid[p] : p;
    cc[3 * k] = pl[4 * k]; cc[3 * k + 1] = pl[4 * k + 1]; cc[3 * k + 2] = pl[4 * k + 2]; dd[k] = pl[4 * k + 3];
  }
  return 0;
}

int tob_separate_self(tob_ctx* c, const double* splines, int n_robots, uint32_t* offsets, double* cc, double* dd, uint64_t cap,
                      uint64_t* total) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (n_robots != c->n_robots()) return fail_msg(c, "tob_separate_self needs all uav_num robots");
  TOB_TRY(stage_splines(c, splines, nullptr, n_robots));
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, 0, n_robots, 1));
  TOB_TRY(pack_self_only(c));
  TOB_TRY(sync_counts(c));
  uint64_t np = c->n_planes;
  if (total) *total = np;
  const int rows = c->rows_all();
  std::vector<uint32_t> off(rows + 1);
  TOB_CUDA(c, cudaMemcpyAsync(off.data(), c->pl_off.p, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  std::vector<double> pl(4 * np);
  const bool want = np && cc && dd && np <= cap;
  if (want) TOB_CUDA(c, cudaMemcpyAsync(pl.data(), c->pl.p, 4 * np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (offsets) for (int r = 0; r <= rows; r++) offsets[r] = off[r];
  if (want)
    for (uint64_t k = 0; k < np; k++) { cc[3 * k] = pl[4 * k]; cc[3 * k + 1] = pl[4 * k + 1]; cc[3 * k + 2] = pl[4 * k + 2]; dd[k] = pl[4 * k + 3]; }
  return 0;
}

int tob_separate_planes(tob_ctx* c, const double* splines, int n_robots, int with_self, uint32_t* offsets, double* cc,
                        double* dd, uint64_t cap, uint64_t* total) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  if (with_self && n_robots != c->n_robots()) return fail_msg(c, "with_self needs all uav_num robots");
  TOB_TRY(stage_splines(c, splines, nullptr, n_robots));
  TOB_TRY(run_checked(c, [&]() { return separate_resident(c, 0, n_robots, with_self); }));
  uint64_t np = c->n_planes;
  if (total) *total = np;
  const int rows = n_robots * c->n_tr;
  std::vector<uint32_t> off(c->rows_all() + 1);
  TOB_CUDA(c, cudaMemcpyAsync(off.data(), c->pl_off.p, off.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  std::vector<double> pl(4 * np);
  if (np && cc && dd && np <= cap)
    TOB_CUDA(c, cudaMemcpyAsync(pl.data(), c->pl.p, 4 * np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (offsets) for (int r = 0; r <= rows; r++) offsets[r] = off[r];
  if (cc && dd && np <= cap)
    for (uint64_t k = 0; k < np; k++) { cc[3 * k] = pl[4 * k]; cc[3 * k + 1] = pl[4 * k + 1]; cc[3 * k + 2] = pl[4 * k + 2]; dd[k] = pl[4 * k + 3]; }
  return 0;
}

int tob_set_planes(tob_ctx* c, int n_robots, const uint32_t* offsets, const double* cc, const double* dd) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (n_robots < 1 || n_robots > c->n_robots()) return fail_msg(c, "tob_set_planes: bad n_robots");
  return pack_planes_from_host(c, 0, n_robots, offsets, cc, dd);
}

// ---- energies ----------------------------------------------------------------------------------------------------------
static int robot_ok(tob_ctx* c, int robot) {
  if (robot < 0 || robot >= c->n_robots()) return fail_msg(c, "robot index out of range");
  return 0;
}

// row_e record of (trial 0, row): up to 8 plane-energy partials (one per 256 planes of the row, barrier.cu) + the bound energy
__global__ void k_plane_energy_only(const double* row_e, const int* bad, const uint32_t* pl_off, int row0, int n_tr, double* out) {
  // single thread: ordered sum like the reference's loop over tr_id
  if (threadIdx.x || blockIdx.x) return;
  double e = 0;
  for (int t = 0; t < n_tr; t++) {
    const uint32_t np = pl_off[row0 + t + 1] - pl_off[row0 + t];
    int V = (int)((np + 255) / 256); V = V < 1 ? 1 : (V > 8 ? 8 : V);
    for (int v = 0; v < V; v++) e += row_e[TOB_EN_REC * (size_t)(row0 + t) + v];
  }
  out[0] = e; out[1] = bad[0];
}

__global__ void k_bound_energy_only(const double* row_e, const int* bad, int row0, int n_tr, double* out) {
  if (threadIdx.x || blockIdx.x) return;
  double e = 0;
  for (int t = 0; t < n_tr; t++) e += row_e[TOB_EN_REC * (size_t)(row0 + t) + (TOB_EN_REC - 1)];
  out[0] = e; out[1] = bad[0];
}

int tob_plane_barrier_energy(tob_ctx* c, int robot, const double* spline, double* e) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  cudaSetDevice(c->device);
  TOB_TRY(upload(c, c->s_spline, spline, 3 * c->T, (size_t)robot * 3 * c->T));
  // a huge piece time switches the bound terms off without touching the plane sum
  double big = 1e300;
  TOB_TRY(upload(c, c->s_ptrial, &big, 1, robot));
  TOB_TRY(energy_trials(c, robot, robot + 1, nullptr, nullptr, c->s_ptrial.p, 1, 0, 1, c->s_e1.p));
  k_plane_energy_only<<<1, 32, 0, c->stream>>>(c->row_e.p, c->row_bad.p + robot, c->pl_off.p, robot * c->n_tr, c->n_tr, c->red.p);
  TOB_LAUNCH_CHECK(c);
  TOB_TRY(read_back(c, c->red.p, 2 * sizeof(double)));
  *e = c->h_pinned[1] != 0 ? INFINITY : c->h_pinned[0];
  return 0;
}

int tob_bound_energy(tob_ctx* c, const double* spline, double piece_time, double* e) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (c->n_planes == 0 && !c->pl_off.p) {   // no plane set yet: install an empty one
    std::vector<uint32_t> off(c->n_tr + 1, 0u);
    TOB_TRY(pack_planes_from_host(c, 0, 1, off.data(), nullptr, nullptr));
  }
  TOB_TRY(upload(c, c->s_spline, spline, 3 * c->T, 0));
  TOB_TRY(upload(c, c->s_ptrial, &piece_time, 1, 0));
  TOB_TRY(energy_trials(c, 0, 1, nullptr, nullptr, c->s_ptrial.p, 1, 0, 1, c->s_e1.p));
  k_bound_energy_only<<<1, 32, 0, c->stream>>>(c->row_e.p, c->row_bad.p, 0, c->n_tr, c->red.p);
  TOB_LAUNCH_CHECK(c);
  TOB_TRY(read_back(c, c->red.p, 2 * sizeof(double)));
  *e = c->h_pinned[1] != 0 ? INFINITY : c->h_pinned[0];
  return 0;
}

int tob_spline_energy(tob_ctx* c, int robot, const tob_state* st, double* e) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  cudaSetDevice(c->device);
  TOB_TRY(put_state(c, robot, st));
  TOB_TRY(energy_trials(c, robot, robot + 1, nullptr, nullptr, c->s_ptime.p, 1, 0, 1, c->s_e1.p));
  TOB_TRY(read_back(c, c->s_e1.p + robot, sizeof(double)));
  *e = c->h_pinned[0];
  return 0;
}

// ---- gradient / direction -------------------------------------------------------------------------------------------------
int tob_piece_blocks(tob_ctx* c, int robot, const tob_state* st, int project_psd, double* g, double* h) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  cudaSetDevice(c->device);
  const int P = c->prm.piece_num;
  TOB_TRY(put_state(c, robot, st));
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, robot, robot + 1, 0));
  TOB_TRY(gradient_blocks(c, robot, robot + 1, project_psd));
  TOB_CUDA(c, cudaMemcpyAsync(g, c->pc_g.p + (size_t)19 * robot * P, (size_t)19 * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(h, c->pc_h.p + (size_t)361 * robot * P, (size_t)361 * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_global_gradient(tob_ctx* c, int robot, const tob_state* st, double* grad, double* hess) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  const int P = c->prm.piece_num, n = 3 * c->T;
  std::vector<double> g((size_t)19 * P), h((size_t)361 * P);
  TOB_TRY(tob_piece_blocks(c, robot, st, 1, g.data(), h.data()));
  // scatter-add of Gradient_admm.h:55-62 (dense output only exists for the unchanged signature)
  const size_t ld = n + 1;
  for (size_t i = 0; i < ld; i++) grad[i] = 0;
  for (size_t i = 0; i < ld * ld; i++) hess[i] = 0;
  for (int sp = 0; sp < P; sp++) {
    const double* g0 = &g[(size_t)19 * sp];
    const double* h0 = &h[(size_t)361 * sp];
    int base = 9 * sp;
    for (int i = 0; i < 18; i++) grad[base + i] += g0[i];
    grad[n] += g0[18];
    for (int j = 0; j < 18; j++)
      for (int i = 0; i < 18; i++) hess[(base + i) + ld * (base + j)] += h0[i + 19 * j];
    hess[n + ld * n] += h0[18 + 19 * 18];
    for (int i = 0; i < 18; i++) {
      hess[(base + i) + ld * n] += h0[i + 19 * 18];
      hess[n + ld * (base + i)] += h0[18 + 19 * i];
    }
  }
  return 0;
}

int tob_descent_direction(tob_ctx* c, int robot, const tob_state* st, int dense_shift, double* direction, double* t_direction,
                          double* wolfe, double* gnorm) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  cudaSetDevice(c->device);
  TOB_TRY(put_state(c, robot, st));
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, robot, robot + 1, 0));
  TOB_TRY(gradient_blocks(c, robot, robot + 1, 1));
  TOB_TRY(solve_directions(c, robot, robot + 1, dense_shift));
  TOB_CUDA(c, cudaMemcpyAsync(direction, c->s_dir.p + (size_t)robot * 3 * c->T, 3 * c->T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(t_direction, c->s_tdir.p + robot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(wolfe, c->s_wolfe.p + robot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(gnorm, c->s_gnorm.p + robot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_TRY(read_back(c, c->solve_status.p + robot, sizeof(int)));
  if (*((int*)c->h_pinned) != 0) return fail_msg(c, "Newton matrix is not positive definite");
  return 0;
}

// ---- CCD steps --------------------------------------------------------------------------------------------------------------
int tob_position_step(tob_ctx* c, const double* spline, const double* direction, double* step) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  TOB_TRY(stage_splines(c, spline, direction, 1));
  TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, 0, 1, 3));   // also resets kmax[0]
  TOB_TRY(ccd_position_steps(c, 0, 1));
  TOB_TRY(read_back(c, c->kmax.p, sizeof(int)));
  int k = *((int*)c->h_pinned);
  double s = 1.0;
  for (int i = 0; i < k; i++) s *= 0.8;
  *step = s;
  return 0;
}

int tob_self_step(tob_ctx* c, const double* splines, const double* directions, int n_robots, int coupled, double* steps) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (n_robots != c->n_robots()) return fail_msg(c, "tob_self_step needs all uav_num robots");
  TOB_TRY(stage_splines(c, splines, directions, n_robots));
  TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, 0, n_robots, 3));
  TOB_TRY(self_ccd_steps(c, coupled, c->s_selfstep.p));
  TOB_CUDA(c, cudaMemcpyAsync(steps, c->s_selfstep.p, (coupled ? 1 : n_robots) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_self_broadphase(tob_ctx* c, const double* P, const double* D, int u, double d, uint32_t* pairs, uint64_t cap,
                        uint64_t* total) {
  // all-pairs box test of one time slot on the host side of the ABI: u <= 64, at most 2016 pairs; the device
  // kernels (k_self_planes / k_self_ccd_filter) apply the same predicate inline.
  if (!c) return 1;
  uint64_t n = 0;
  std::vector<double> lo(3 * u), hi(3 * u);
  for (int i = 0; i < u; i++)
    for (int k = 0; k < 3; k++) {
      double l = INFINITY, h = -INFINITY;
      for (int j = 0; j < 6; j++) {
        double v = P[(size_t)18 * i + 6 * k + j];
        if (v < l) l = v; if (v > h) h = v;
        if (D) { double w = v + D[(size_t)18 * i + 6 * k + j]; if (w < l) l = w; if (w > h) h = w; }
      }
      lo[3 * i + k] = l; hi[3 * i + k] = h;
    }
  for (int a = 0; a < u; a++)
    for (int b = a + 1; b < u; b++) {
      bool hit = true;
      for (int k = 0; k < 3; k++)
        if (hi[3 * a + k] + d < lo[3 * b + k] || lo[3 * a + k] > hi[3 * b + k] + d) hit = false;
      if (hit) { if (n < cap && pairs) { pairs[2 * n] = a; pairs[2 * n + 1] = b; } n++; }
    }
  if (total) *total = n;
  return 0;
}

// ---- line search -------------------------------------------------------------------------------------------------------------------
int tob_line_search(tob_ctx* c, int robot, tob_state* st, const double* direction, double t_direction, double wolfe, double* step_io) {
  TOB_TRY(need(c, true, false)); TOB_TRY(robot_ok(c, robot));
  cudaSetDevice(c->device);
  if (!step_io) return fail_msg(c, "tob_line_search: step_io is required");
  const int T = c->T;
  TOB_TRY(put_state(c, robot, st));
  TOB_TRY(upload(c, c->s_dir, direction, 3 * T, (size_t)robot * 3 * T));
  TOB_TRY(upload(c, c->s_tdir, &t_direction, 1, robot));
  TOB_TRY(upload(c, c->s_wolfe, &wolfe, 1, robot));
  k_fill_int<<<div_up(c->n_robots(), 64), 64, 0, c->stream>>>(c->kmax.p, c->n_robots(), 0);
  TOB_LAUNCH_CHECK(c);
  int use_self = 0;
  ls_policy(c, robot, robot + 1, false);
  if (*step_io < 0) {           // Optimization3D_admm::spline_line_search: the bound is Step::position_step
    if (c->n_pts == 0) return fail_msg(c, "tob_line_search: no point cloud");
    TOB_TRY(compute_rows(c, c->s_spline.p, c->s_dir.p, nullptr, robot, robot + 1, 3));
    TOB_TRY(ccd_position_steps(c, robot, robot + 1));
  } else {                      // Optimization3D_multi::spline_line_search: the caller's bound (<= 1)
    if (*step_io > 1.0) return fail_msg(c, "tob_line_search: a caller-provided step bound must be <= 1");
    TOB_TRY(upload(c, c->s_selfstep, step_io, 1, robot));
    use_self = 1;
  }
  TOB_TRY(line_search_begin(c, robot, robot + 1));
  k_ls_init<<<1, 64, 0, c->stream>>>(robot, robot + 1, c->kmax.p, c->d_steps.p, c->s_selfstep.p, use_self, c->s_ptime.p, c->s_tdir.p,
                                     c->s_step.p, c->s_tstep.p, c->s_ttime.p, c->s_done.p, c->dc.p, c->ls_rounds, c->row_bad.p, nullptr);
  TOB_LAUNCH_CHECK(c);
  c->ls_e0_ready = false;            // function-level call: the starting point is evaluated by the energy kernel
  TOB_TRY(line_search(c, robot, robot + 1, -1));
  TOB_CUDA(c, cudaMemcpyAsync(st->spline, c->s_spline.p + (size_t)robot * 3 * T, 3 * T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(st->piece_time, c->s_ptime.p + robot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TOB_TRY(read_back(c, c->s_step.p + robot, sizeof(double)));
  *step_io = c->h_pinned[0];
  return 0;
}

int tob_last_wolfe(tob_ctx* c, double* wolfe) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (!c->states_valid || !wolfe) return fail_msg(c, "tob_last_wolfe: no iteration has run");
  // the reference's global `wolfe` after an iteration is the value left by update_slack_lambda's last piece; callers only
  // consume the one written by the direction solve, which is what is kept here (last robot, Optimization3D_multi.h:730)
  TOB_TRY(read_back(c, c->s_wolfe.p + (c->n_robots() - 1), sizeof(double)));
  *wolfe = c->h_pinned[0];
  return 0;
}

int tob_row_blocks(tob_ctx* c, const double* spline, double piece_time, int tr_id, int which, double* grad18, double* hess324,
                   double* g_t, double* h_t, double* partgrad18) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (tr_id < 0 || tr_id >= c->n_tr) return fail_msg(c, "tob_row_blocks: tr_id out of range");
  if (!c->pl_off.p) {
    std::vector<uint32_t> off(c->n_tr + 1, 0u);
    TOB_TRY(pack_planes_from_host(c, 0, 1, off.data(), nullptr, nullptr));
  }
  TOB_TRY(upload(c, c->s_spline, spline, 3 * c->T, 0));
  TOB_TRY(upload(c, c->s_ptime, &piece_time, 1, 0));
  c->states_valid = false;
  TOB_TRY(compute_rows(c, c->s_spline.p, nullptr, nullptr, 0, 1, 0));
  TOB_CUDA(c, c->scratch2.ensure(362));
  TOB_TRY(row_blocks(c, tr_id, which, c->scratch2.p));
  TOB_TRY(read_back(c, c->scratch2.p, 362 * sizeof(double)));
  const double* h = c->h_pinned;
  if (grad18) memcpy(grad18, h, 18 * sizeof(double));
  if (hess324) memcpy(hess324, h + 18, 324 * sizeof(double));
  if (g_t) *g_t = h[342];
  if (h_t) *h_t = h[343];
  if (partgrad18) memcpy(partgrad18, h + 344, 18 * sizeof(double));
  return 0;
}

// ---- slack / dual ---------------------------------------------------------------------------------------------------------------
int tob_update_slack_lambda(tob_ctx* c, tob_state* st) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  TOB_TRY(put_state(c, 0, st));
  TOB_TRY(slack_update(c, 0, 1));
  const int P = c->prm.piece_num;
  cudaStream_t s = c->stream;   // spline and piece time are inputs only: not written back
  TOB_CUDA(c, cudaMemcpyAsync(st->p_slack, c->s_pslack.p, 18 * P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->t_slack, c->s_tslack.p, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->p_lambda, c->s_plambda.p, 18 * P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaMemcpyAsync(st->t_lambda, c->s_tlambda.p, P * sizeof(double), cudaMemcpyDeviceToHost, s));
  TOB_CUDA(c, cudaStreamSynchronize(s));
  return 0;
}

int tob_slack_terms(tob_ctx* c, const double* c_spline, double piece_time, const double* p_part, double t_part,
                    const double* p_lambda, double t_lambda, int consensus, double* energy, double* grad19, double* hess361) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (!p_part) return fail_msg(c, "tob_slack_terms: p_part is required");
  double in[57];
  for (int i = 0; i < 18; i++) { in[i] = c_spline ? c_spline[i] : 0.0; in[18 + i] = p_part[i]; in[36 + i] = p_lambda ? p_lambda[i] : 0.0; }
  in[54] = piece_time; in[55] = t_part; in[56] = t_lambda;
  TOB_CUDA(c, c->scratch.ensure(57)); TOB_CUDA(c, c->scratch2.ensure(381));
  TOB_TRY(upload(c, c->scratch, in, 57));
  TOB_TRY(slack_terms(c, c->scratch.p, consensus, c->scratch2.p));
  TOB_TRY(read_back(c, c->scratch2.p, 381 * sizeof(double)));
  if (energy) *energy = c->h_pinned[0];
  if (grad19) memcpy(grad19, c->h_pinned + 1, 19 * sizeof(double));
  if (hess361) memcpy(hess361, c->h_pinned + 20, 361 * sizeof(double));
  return 0;
}

// ---- iteration ---------------------------------------------------------------------------------------------------------------------
// all robots' states move in 6 transfers (one per array type) through a pinned staging area instead of 6 per robot
static int stage_ensure(tob_ctx* c, size_t doubles) {
  if (doubles <= c->h_stage_cap) return 0;
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->h_stage = nullptr; c->h_stage_cap = 0;
  TOB_CUDA(c, cudaMallocHost((void**)&c->h_stage, doubles * sizeof(double)));
  c->h_stage_cap = doubles;
  return 0;
}

int tob_states_upload(tob_ctx* c, const tob_state* states, int n_robots) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (n_robots != c->n_robots()) return fail_msg(c, "tob_states_upload: n_robots must equal uav_num");
  const size_t U = n_robots, P = c->prm.piece_num, T = c->T;
  const size_t n_sp = 3 * T, n_ps = 18 * P, n_ts = P;
  TOB_TRY(stage_ensure(c, U * (n_sp + 1 + 2 * n_ps + 2 * n_ts)));
  double* h = c->h_stage;
  double *h_sp = h, *h_pt = h_sp + U * n_sp, *h_ps = h_pt + U, *h_ts = h_ps + U * n_ps, *h_pl = h_ts + U * n_ts, *h_tl = h_pl + U * n_ps;
  for (size_t u = 0; u < U; u++) {
    memcpy(h_sp + u * n_sp, states[u].spline, n_sp * sizeof(double));
    h_pt[u] = *states[u].piece_time;
    memcpy(h_ps + u * n_ps, states[u].p_slack, n_ps * sizeof(double));
    memcpy(h_ts + u * n_ts, states[u].t_slack, n_ts * sizeof(double));
    memcpy(h_pl + u * n_ps, states[u].p_lambda, n_ps * sizeof(double));
    memcpy(h_tl + u * n_ts, states[u].t_lambda, n_ts * sizeof(double));
  }
  TOB_TRY(upload(c, c->s_spline, h_sp, U * n_sp)); TOB_TRY(upload(c, c->s_ptime, h_pt, U));
  TOB_TRY(upload(c, c->s_pslack, h_ps, U * n_ps)); TOB_TRY(upload(c, c->s_tslack, h_ts, U * n_ts));
  TOB_TRY(upload(c, c->s_plambda, h_pl, U * n_ps)); TOB_TRY(upload(c, c->s_tlambda, h_tl, U * n_ts));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->states_valid = true;
  return 0;
}

int tob_states_download(tob_ctx* c, tob_state* states, int n_robots) {
  TOB_TRY(need(c, true, false));
  cudaSetDevice(c->device);
  if (n_robots != c->n_robots()) return fail_msg(c, "tob_states_download: n_robots must equal uav_num");
  const size_t U = n_robots, P = c->prm.piece_num, T = c->T;
  const size_t n_sp = 3 * T, n_ps = 18 * P, n_ts = P;
  TOB_TRY(stage_ensure(c, U * (n_sp + 1 + 2 * n_ps + 2 * n_ts)));
  double* h = c->h_stage;
  double *h_sp = h, *h_pt = h_sp + U * n_sp, *h_ps = h_pt + U, *h_ts = h_ps + U * n_ps, *h_pl = h_ts + U * n_ts, *h_tl = h_pl + U * n_ps;
  cudaStream_t st = c->stream;
  TOB_CUDA(c, cudaMemcpyAsync(h_sp, c->s_spline.p, U * n_sp * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaMemcpyAsync(h_pt, c->s_ptime.p, U * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaMemcpyAsync(h_ps, c->s_pslack.p, U * n_ps * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaMemcpyAsync(h_ts, c->s_tslack.p, U * n_ts * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaMemcpyAsync(h_pl, c->s_plambda.p, U * n_ps * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaMemcpyAsync(h_tl, c->s_tlambda.p, U * n_ts * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  // a sharded context returns the robots it owns (the other slots of `states` are left untouched: their owners write them)
  const size_t ub = c->sharded() ? (size_t)c->own_begin : 0, ue = c->sharded() ? (size_t)c->own_end : U;
  for (size_t u = ub; u < ue; u++) {
    memcpy(states[u].spline, h_sp + u * n_sp, n_sp * sizeof(double));
    *states[u].piece_time = h_pt[u];
    memcpy(states[u].p_slack, h_ps + u * n_ps, n_ps * sizeof(double));
    memcpy(states[u].t_slack, h_ts + u * n_ts, n_ts * sizeof(double));
    memcpy(states[u].p_lambda, h_pl + u * n_ps, n_ps * sizeof(double));
    memcpy(states[u].t_lambda, h_tl + u * n_ts, n_ts * sizeof(double));
  }
  return 0;
}

int tob_admm_iterate(tob_ctx* c, int iters, int mode, double* gnorm) {
  TOB_TRY(need(c, true, true));
  cudaSetDevice(c->device);
  if (!c->states_valid) return fail_msg(c, "tob_admm_iterate: call tob_states_upload first");
  TOB_TRY(clear_overflow(c));
  for (int i = 0; i < iters; i++) TOB_TRY(iterate_once(c, mode, (i == iters - 1) ? gnorm : nullptr));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int tob_optimization(tob_ctx* c, tob_state* states, int n_robots, int mode, double* gnorm) {
  TOB_TRY(tob_states_upload(c, states, n_robots));
  TOB_TRY(tob_admm_iterate(c, 1, mode, gnorm));
  return tob_states_download(c, states, n_robots);
}

int tob_get_counters(const tob_ctx* cc, tob_counters* out) {
  if (!cc || !out) return 1;
  tob_ctx* c = const_cast<tob_ctx*>(cc);
  cudaSetDevice(c->device);
  TOB_TRY(sync_counts(c));     // the pair counters are accumulated on the device
  *out = c->ctr;
  out->dcd_candidates = c->h_dc->dcd_candidates;
  out->planes = c->h_dc->planes;
  out->ccd_candidates = c->h_dc->ccd_candidates;
  out->energy_plane_evals = c->h_dc->energy_plane_evals;
  out->barrier_terms = c->h_dc->barrier_terms;
  out->live_planes = c->h_dc->n_live;
  out->refine_capped = c->h_dc->opt_capped;
  out->np_kdop_groups = c->h_dc->np_kdop_groups;
  out->np_gjk_iters = c->h_dc->np_gjk_iters;
  out->ccd_gjk_iters = c->h_dc->ccd_gjk_iters;
  out->ccd_kdop_pass = c->h_dc->ccd_kdop_pass;
  out->np_kdop_exact = c->h_dc->np_kdop_exact;
  out->np_band = c->h_dc->np_band;
  for (int i = 0; i < 8; i++) out->ls_rung_hist[i] = c->h_dc->ls_hist[i];
  out->ls_rungs_skipped = c->h_dc->ls_skipped;
  return 0;
}
int tob_reset_counters(tob_ctx* c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  memset(&c->ctr, 0, sizeof(c->ctr));
  TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->dcd_candidates, 0, 5 * sizeof(unsigned long long), c->stream));
  TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->np_kdop_groups, 0, 15 * sizeof(unsigned long long), c->stream));
  return 0;
}

static const char* kKernelNames[K_COUNT] = {"k_rows", "k_bp_count", "k_bp_fill", "k_bp_top+k_np_top", "k_narrow",
                                            "k_pack", "k_self_planes", "k_row_energy", "k_robot_ls", "k_row_grad",
                                            "k_piece", "k_solve_bcr", "k_bp_ccd", "k_self_ccd_filter", "k_slack", "misc",
                                            "k_live_refine"};

int tob_profile_enable(tob_ctx* c, int on) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  prof_collect(c);
  c->prof_on = on != 0;
  if (on) for (int i = 0; i < 32; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
  return 0;
}

int tob_profile_read(tob_ctx* c, int kid, double* ms_total, uint64_t* launches, const char** name) {
  if (!c || kid < 0 || kid >= K_COUNT) return 1;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  prof_collect(c);
  if (ms_total) *ms_total = c->prof_ms[kid];
  if (launches) *launches = c->prof_n[kid];
  if (name) *name = kKernelNames[kid];
  return 0;
}

int tob_set_shard(tob_ctx* c, int first, int count, int n_total, tob_allgather_fn ag, tob_allreduce_fn ar, void* user) {
  TOB_TRY(need(c, false, false));
  if (n_total != c->n_robots() || first < 0 || count < 1 || first + count > n_total) return fail_msg(c, "tob_set_shard: bad range");
  if (c->nccl_comm) return fail_msg(c, "tob_set_shard: a NCCL communicator is attached (tob_nccl_detach first)");
  if (ag && (n_total % count || first % count)) return fail_msg(c, "tob_set_shard: the callback exchange needs equal blocks (n_total divisible by count)");
  c->own_begin = first; c->own_end = first + count; c->ag = ag; c->ar = ar; c->cb_user = user;
  c->comm_world = ag ? n_total / count : 1; c->comm_rank = ag ? first / count : 0;
  TOB_CUDA(c, c->ovf_all.ensure((size_t)c->comm_world + 1));
  graph_drop(c);
  return 0;
}

void* tob_stream(tob_ctx* c) { return c ? (void*)c->stream : nullptr; }

int tob_build_stats(const tob_ctx* c, double* ms, uint64_t* points) {
  if (!c) return 1;
  if (ms) *ms = c->build_ms;
  if (points) *points = c->build_points;
  return 0;
}

// ---- FP64 pipe peak ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, cc = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int tob_fp64_peak(tob_ctx* c, double* tflops) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  const int blocks = c->sm_count * 8, threads = 256, iters = 1 << 14;
  TOB_CUDA(c, c->scratch.ensure((size_t)blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<<<blocks, threads, 0, c->stream>>>(c->scratch.p, iters);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0, c->stream);
    k_dfma<<<blocks, threads, 0, c->stream>>>(c->scratch.p, iters);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double flops = 2.0 * 8 * (double)iters * blocks * threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
