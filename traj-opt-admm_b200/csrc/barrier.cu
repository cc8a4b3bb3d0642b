// barrier.cu -- FP64 log-barrier energy, gradient and Hessian-block reductions over (sub-segment x plane) pairs.
//
// Replaces Energy_admm::{plane_barrier_energy :46-96, bound_energy :98-170, spline_energy :16-44} and
// Gradient_admm::{local_plane_barrier_gradient :331-407, local_bound_gradient :409-572, local_spline_gradient :67-164,
// the per-piece PSD projection of global_spline_gradient :13-65}.
//
// Data flow (all per row = robot x sub-segment, planes packed CSR by row, 32 B per plane: cx,cy,cz,d):
//   k_row_energy   one CTA per row: sum_k sum_{j<6} b(P_j.c_k + d_k), + the 9 velocity/acceleration bound terms
//   k_robot_energy one CTA per robot: ordered sum over its rows + ADMM consensus terms -> E (inf if any d<=0)
//   k_row_grad     one CTA per row: the 18x18 rank-1 updates of the reference collapse algebraically to
//                  H_row = sum_j (b_j b_j^T) (x) S_j with S_j = sum_k e2_jk c_k c_k^T (3x3), g_row = sum_j b_j (x) g_j;
//                  the bound terms have the same shape (a a^T) (x) M3.  So a row is summarised by 15 "terms"
//                  (6 plane + 5 velocity + 4 acceleration), 12 doubles each, independent of the plane count.
//   k_piece        one CTA per (robot, piece): expands the 8x15 terms into the 19x19 block, adds the consensus
//                  terms, Cholesky test, eigen-shift when not SPD.
// Reductions are fixed-order (shuffle tree + ordered cross-warp sum): bitwise reproducible run to run.
// Tolerance vs the reference: summation order differs, FMA contraction allowed here -> ~1e-13 relative.
#include <stdlib.h>

#include "ctx.cuh"
#include "dense.cuh"
#include "fastlog.cuh"

namespace tob {

#define ROW_TERMS 15
#define TERM_SZ 12                 // M3: xx xy xz yy yz zz | g3 | pg3
#define ROW_REC (ROW_TERMS * TERM_SZ + 2)

// table of tob_log_pos (fastlog.cuh); the barrier kernels copy it to shared memory
__device__ const LogTabEntry g_logtab[TOB_LOGTAB_N] = {
#include "logtab.inc"
};
__device__ __forceinline__ void load_logtab(LogTabEntry* s_lt) {
  for (int i = threadIdx.x; i < TOB_LOGTAB_N; i += blockDim.x) s_lt[i] = g_logtab[i];
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- energy --------------------------------------------------------------------------------------------------------
// One launch evaluates a BATCH of line-search trial points: trial k of robot u is spline + tstep[u][k]*dir (element-wise,
// then P = basis*bz with the reference's unfused left-to-right arithmetic: explicit __dmul_rn/__dadd_rn because this file
// is compiled with FMA contraction on), trial piece time = ttime[u][k].
//
// Summation structure (what makes the result independent of the launch shape, of the batch a problem sits in and of the
// sharding): the planes of a row are dealt in chunks of 32 to V(row) = clamp(ceil(planes / 256), 1, 8) VIRTUAL WARPS
// (chunk i belongs to virtual warp i mod V); a virtual warp sums its terms lane-wise in chunk order, reduces with the
// xor-shuffle tree and stores ONE partial; the robot kernels add the V partials of a row in order.  V depends on the row's
// plane count only.  How many physical warps / CTAs execute the virtual warps of a row is a free scheduling choice (VL).
//
// Early exit (Energy_admm.h:81-82: the reference returns INFINITY at the first d <= 0): a warp that meets a violated plane
// raises the flag of its (robot, trial) and stops; every other warp of that trial -- of any row of the robot -- polls the
// flag and stops too.  An infeasible trial costs a few plane chunks instead of a full pass.
// In-band compaction: only terms inside the barrier band (0 < d < margin) need a logarithm; they are queued per warp
// (ballot + prefix, deterministic order) and evaluated on dense lanes.
#define EN_VMAX 8
#define EN_VPLANES 256
#define EN_REC (EN_VMAX + 1)        // per (trial, row): EN_VMAX plane-energy partials + the bound energy
__host__ __device__ __forceinline__ int en_vwarps(uint32_t np) {
  const int v = (int)((np + EN_VPLANES - 1) / EN_VPLANES);
  return v < 1 ? 1 : (v > EN_VMAX ? EN_VMAX : v);
}

struct EnergyArgs {
  const double *spline, *dir;   // robots x 3T ; dir may be null (trial == current point)
  const double *tstep, *ttime;  // robots x KT ; tstep may be null (=0)
  const double* basis;          // n_tr x 36
  const double* pl;             // planes x 4
  const uint32_t* pl_off;       // rows+1
  const double* weight;         // n_tr
  double margin, vel_limit, acc_limit;
  int n_tr, res, T, row_begin, n_rows, rows_all, KT, k0, nk;
  const uint32_t* items;        // (row << 3 | v) of every virtual warp v >= 1 of the context's rows (k_en_items), any order
  double* row_e;                // KT x rows_all x EN_REC
  int* bad;                     // robots x KT: trial infeasible (some d <= 0); raised here, cleared by the caller
  const int* done;              // per robot, may be null: robots whose line search has finished are skipped
  const int* pending;           // may be null; else the robots still backtracking after the previous Armijo round: 0 = this
                                // round (launched ahead of the host, inside the graph) has nothing to do and every CTA leaves
  DevCounts* dc;                // barrier_terms counter, n_en_items
};

// Every row with more than EN_VPLANES planes lists its virtual warps 1 .. V-1 here, once per plane set (after the CSR is
// built): the energy launches then run one CTA per row for virtual warp 0 (+ the bound terms) and a fixed number of extra
// CTAs that share the listed virtual warps, instead of V CTAs per row of which nearly all would find nothing to do.
__global__ void __launch_bounds__(256) k_en_items(const uint32_t* __restrict__ pl_off, int rows_all, uint32_t* __restrict__ items,
                                                  uint32_t* __restrict__ item_base, DevCounts* dc) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows_all || dc->overflow) return;
  const int V = en_vwarps(pl_off[row + 1] - pl_off[row]);
  if (V > 1) {
    const uint32_t base = atomicAdd(&dc->n_en_items, (uint32_t)(V - 1));
    item_base[row] = base;         // item of (row, v) sits at base + v - 1: where the gradient pass leaves its partial sums
    for (int v = 1; v < V; v++) items[base + v - 1] = ((uint32_t)row << 3) | (uint32_t)(v - 1);
  }
}

#define EN_MAXT TOB_LS_TRIALS
// one warp: virtual warp v of `row` for trial k (and, when v == 0, the bound terms of the row)
// CP_SHARED: the 18 control-point coordinates are read from shared memory in every chunk (uniform address: one broadcast per
// load) instead of living in 36 registers -- the variants that trade registers for resident warps
template <bool CP_SHARED>
__device__ __forceinline__ void en_virtual_warp(const EnergyArgs& a, int row, int v, int k, double* sP, double* sBz, double* q,
                                                const LogTabEntry* s_lt, unsigned& n_act, unsigned& n_pl) {
  const int lane = threadIdx.x & 31;
  const int robot = row / a.n_tr, tr = row - robot * a.n_tr;
  // every word that decides whether there is anything to do is loaded before the first test: one memory round trip
  const uint32_t p0 = a.pl_off[row], p1 = a.pl_off[row + 1];
  const int is_done = a.done ? a.done[robot] : 0;
  const int seen0 = a.bad[robot * a.KT + k];    // plain (cached) load: a stale 0 only costs the work the flag would have saved
  if (is_done || seen0) return;    // (an infeasible trial's energy is +inf whatever this row adds)
  const int V = en_vwarps(p1 - p0);
  volatile int* flag = a.bad + robot * a.KT + k;
  double* out = a.row_e + ((size_t)k * a.rows_all + row) * EN_REC;
  if (lane < 18) {
    const int mm = lane % 6, ax = lane / 6;
    const size_t gi = (size_t)robot * 3 * a.T + (size_t)ax * a.T + 3 * (tr / a.res) + mm;
    double x = a.spline[gi];
    if (a.dir && a.tstep) x = __dadd_rn(x, __dmul_rn(a.tstep[robot * a.KT + k], a.dir[gi]));
    sBz[lane] = x;
  }
  __syncwarp();
  if (lane < 18) {
    const int j = lane % 6, ax = lane / 6;
    const double* B = a.basis + (size_t)36 * tr;
    double acc = 0;
    for (int i = 0; i < 6; i++) acc = __dadd_rn(acc, __dmul_rn(B[j + 6 * i], sBz[i + 6 * ax]));
    sP[lane] = acc;
  }
  __syncwarp();
  const double w = a.weight[tr], m = a.margin, inv_m = 1.0 / m;
  double cp[18];
  const volatile double* vP = sP;
  if (!CP_SHARED) {
#pragma unroll
    for (int i = 0; i < 18; i++) cp[i] = sP[i];
  }
  bool stop = false;
  {
    double e = 0;
    uint32_t p = p0 + (uint32_t)v * 32u + lane;
    const uint32_t stride = 32u * (uint32_t)V;
    double4 nxt = make_double4(0, 0, 0, 0);
    if (p < p1) nxt = *reinterpret_cast<const double4*>(a.pl + (size_t)4 * p);
    uint32_t trip = 0;
    for (uint32_t base = p0 + (uint32_t)v * 32u; base < p1; base += stride, p += stride, trip++) {
      const double4 pl = nxt;
      const bool have = p < p1;
      if (p + stride < p1) nxt = *reinterpret_cast<const double4*>(a.pl + (size_t)4 * (p + stride));
      // the flag of the trial is polled every eighth chunk (an L2 round trip on a word that every warp of the trial reads):
      // rows of fewer than 8 chunks per virtual warp never poll; the load is issued here and consumed after the distances
      const int seen = (trip & 7u) == 7u ? *flag : 0;
      double d[6];
      bool bad = false;
      unsigned cnt = 0;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (CP_SHARED) d[j] = vP[j] * pl.x + vP[j + 6] * pl.y + vP[j + 12] * pl.z + pl.w;
        else d[j] = cp[j] * pl.x + cp[j + 6] * pl.y + cp[j + 12] * pl.z + pl.w;
        bad |= have && (d[j] <= 0);
        const bool act = have && d[j] > 0 && d[j] < m;
        const unsigned bm = __ballot_sync(0xffffffffu, act);
        if (act) q[cnt + __popc(bm & ((1u << lane) - 1u))] = d[j];
        cnt += __popc(bm);             // uniform
      }
      n_pl += have;
      if (__any_sync(0xffffffffu, bad) || seen) { stop = true; if (lane == 0) *flag = 1; break; }
      __syncwarp();
      // the queued terms on dense lanes, three independent logarithm chains at a time (an FP64 result takes ~40 cycles here:
      // one chain after the other would cost a chunk 3x the latency; six at a time cost registers, i.e. resident warps);
      // a slot beyond the queue evaluates log(1) * 0
      for (unsigned t0 = lane; t0 < cnt + lane; t0 += 96u) {      // uniform trip count: ceil(cnt / 96)
        const unsigned t1 = t0 + 32u, t2 = t0 + 64u;
        const double d0 = t0 < cnt ? q[t0] : m, d1 = t1 < cnt ? q[t1] : m, d2 = t2 < cnt ? q[t2] : m;
        const double m0 = d0 - m, m1 = d1 - m, m2 = d2 - m;
        const double l0 = tob_log_pos(d0 * inv_m, s_lt), l1 = tob_log_pos(d1 * inv_m, s_lt), l2 = tob_log_pos(d2 * inv_m, s_lt);
        e += ((m0 * m0) * l0 + (m1 * m1) * l1) + (m2 * m2) * l2;
      }
      n_act += cnt;                    // uniform value: counted once per warp by the caller
      __syncwarp();
    }
    if (!stop) {
      e = warp_sum(e) * -w;
      if (lane == 0) out[v] = e;
    }
  }
  // bound terms: virtual warp 0 only; lanes 0..4 velocity, 5..8 acceleration
  if (v == 0 && !stop) {
    double eb = 0;
    int bad = 0;
    if (lane < 9) {
      const double t = a.ttime[robot * a.KT + k];
      double d;
      if (lane < 5) {
        int j = lane;
        double vx = 5 * (sP[j + 1] - sP[j]), vy = 5 * (sP[j + 7] - sP[j + 6]), vz = 5 * (sP[j + 13] - sP[j + 12]);
        d = a.vel_limit - sqrt(vx * vx + vy * vy + vz * vz) / (w * t);
      } else {
        int j = lane - 5;
        double ax = 20 * (sP[j + 2] - 2 * sP[j + 1] + sP[j]), ay = 20 * (sP[j + 8] - 2 * sP[j + 7] + sP[j + 6]),
               az = 20 * (sP[j + 14] - 2 * sP[j + 13] + sP[j + 12]);
        d = a.acc_limit - sqrt(ax * ax + ay * ay + az * az) / (w * w * t * t);
      }
      if (d <= 0) bad = 1;
      else if (d < m) eb = -w * (d - m) * (d - m) * log(d / m);
    }
    eb = warp_sum(eb);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      out[EN_VMAX] = eb;
      if (bad) *flag = 1;
    }
  }
  __syncwarp();
}

// grid = n_rows + extra CTAs; one warp per trial point of the launch (blockDim = 32 * nk).  CTA b < n_rows: virtual warp 0
// of row b.  The extra CTAs share the listed virtual warps (v >= 1) of the heavy rows.
// MAXT = most trial points (warps) of a launch of this variant, MINB = resident CTAs per SM its registers are allotted for:
// the kernel waits on plane loads (ncu: long scoreboard, 20 % of the warp slots in use with 106 registers), so the variants
// for the short launches of the many-row line-search policy trade registers for resident warps.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(32 * MAXT, MINB) k_row_energy(EnergyArgs a) {
  constexpr bool CPS = MINB > 1;
  const int lane = threadIdx.x & 31, kk = threadIdx.x >> 5, k = a.k0 + kk;
  __shared__ double sPall[MAXT][18], sBzall[MAXT][18];
  __shared__ double sQ[MAXT][192];          // in-band terms of one chunk (32 planes x 6 control points)
  __shared__ LogTabEntry s_lt[TOB_LOGTAB_N];
  if (a.dc->overflow) return;      // the plane CSR of this iteration was not built: the host grows the buffers and retries
  if (a.pending && *a.pending == 0) return;
  load_logtab(s_lt);
  __syncthreads();
  unsigned n_act = 0, n_pl = 0;
  if ((int)blockIdx.x < a.n_rows) {
    en_virtual_warp<CPS>(a, a.row_begin + (int)blockIdx.x, 0, k, sPall[kk], sBzall[kk], sQ[kk], s_lt, n_act, n_pl);
  } else {
    const uint32_t n_items = a.dc->n_en_items, extra = gridDim.x - (uint32_t)a.n_rows;
    for (uint32_t idx = blockIdx.x - (uint32_t)a.n_rows; idx < n_items; idx += extra) {
      const uint32_t it = a.items[idx];
      const int row = (int)(it >> 3), v = (int)(it & 7u) + 1;
      if (row < a.row_begin || row >= a.row_begin + a.n_rows) continue;
      en_virtual_warp<CPS>(a, row, v, k, sPall[kk], sBzall[kk], sQ[kk], s_lt, n_act, n_pl);
    }
  }
  n_pl = __reduce_add_sync(0xffffffffu, n_pl);
  if (lane == 0) {
    if (n_act) atomicAdd(&a.dc->barrier_terms, (unsigned long long)n_act);
    if (n_pl) atomicAdd(&a.dc->energy_plane_evals, (unsigned long long)n_pl);   // planes x trials really evaluated
  }
}

// variant by the number of trial points of the launch; many rows (throughput regime): the occupancy variants
static void launch_row_energy(tob_ctx* c, const EnergyArgs& a, int grid, int nk, int nrows) {
  const char* e = getenv("TRAJOPT_B200_EN_OCC");       // 0: one variant for everything (A/B measurements, tests)
  const int occ = e ? atoi(e) : 1;
  const bool many = occ && nrows >= 8192;
  if (many && nk == 1) k_row_energy<1, 28><<<grid, 32, 0, c->stream>>>(a);
  else if (many && nk == 2) k_row_energy<2, 14><<<grid, 64, 0, c->stream>>>(a);
  else if (many && nk <= 4 && occ == 2) k_row_energy<4, 8><<<grid, 32 * nk, 0, c->stream>>>(a);
  else if (many && nk <= 4) k_row_energy<4, 6><<<grid, 32 * nk, 0, c->stream>>>(a);
  else k_row_energy<EN_MAXT, 1><<<grid, 32 * nk, 0, c->stream>>>(a);
}

// plane-barrier energy of one (trial, row): its V partials in order
__device__ __forceinline__ double row_plane_sum(const double* row_e, size_t o, const uint32_t* pl_off, int row) {
  const double* p = row_e + o * EN_REC;
  const int V = en_vwarps(pl_off[row + 1] - pl_off[row]);
  double s = p[0];
  for (int v = 1; v < V; v++) s += p[v];
  return s;
}

struct RobotEnergyArgs {
  const double *spline, *dir, *tstep, *ttime;
  const double *pslack, *tslack, *plambda, *tlambda, *convert;
  const double* row_e;
  const uint32_t* pl_off;
  int* bad;                                     // robots x KT (raised by k_row_energy)
  double lambda, mu;
  int n_tr, P, T, robot_begin, rows_all, KT, k0;
  double* e_out;                                // robots x KT
};

__global__ void __launch_bounds__(128) k_robot_energy(RobotEnergyArgs a) {
  const int robot = a.robot_begin + blockIdx.x, k = a.k0 + blockIdx.y;
  __shared__ double s_part[4];
  const bool infeasible = a.bad[robot * a.KT + k] != 0;      // the row sums of an infeasible trial are incomplete: unused
  double e = 0;
  for (int tr = threadIdx.x; tr < a.n_tr && !infeasible; tr += blockDim.x) {
    const int row = robot * a.n_tr + tr;
    size_t o = (size_t)k * a.rows_all + row;
    e += a.lambda * row_plane_sum(a.row_e, o, a.pl_off, row) + a.lambda * a.row_e[o * EN_REC + EN_VMAX];
  }
  // consensus terms, one thread per piece
  const double t = a.ttime[robot * a.KT + k];
  const double st = (a.dir && a.tstep) ? a.tstep[robot * a.KT + k] : 0.0;
  for (int sp = threadIdx.x; sp < a.P; sp += blockDim.x) {
    const double* C = a.convert + (size_t)36 * sp;
    double acc = 0;
    for (int ax = 0; ax < 3; ax++) {
      double bz[6];
      for (int q = 0; q < 6; q++) {
        size_t g = (size_t)robot * 3 * a.T + (size_t)ax * a.T + 3 * sp + q;
        bz[q] = a.dir ? __dadd_rn(a.spline[g], __dmul_rn(st, a.dir[g])) : a.spline[g];
      }
      for (int r = 0; r < 6; r++) {
        double cx = 0;
        for (int q = 0; q < 6; q++) cx += C[r + 6 * q] * bz[q];
        size_t s = (size_t)robot * 18 * a.P + (size_t)ax * 6 * a.P + 6 * sp + r;
        double pd = cx - a.pslack[s];
        acc += a.mu / 2.0 * pd * pd + a.plambda[s] * pd;
      }
    }
    double dt = t - a.tslack[(size_t)robot * a.P + sp];
    acc += a.mu / 2.0 * dt * dt + a.tlambda[(size_t)robot * a.P + sp] * dt;
    e += acc;
  }
  e = warp_sum(e);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) s_part[wp] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
    a.e_out[robot * a.KT + k] = infeasible ? INFINITY : tot;
  }
}

// Decoupled line search, one CTA per robot: the robot sums of k_robot_energy for trials k0..KT-1 (same summation
// order), then thread 0 takes the Armijo decision in ladder order exactly like the reference's
//   while(e - 1e-4*wolfe*step < E(x + step*dir)) step *= 0.8     (Optimization3D_admm.h:537-544)
// : the first rung that fails the while-condition is accepted; if all fail, the next 8 rungs are laid out and the
// robot is counted in dc->ls_pending[slot].  wolfe_idx < 0: own wolfe, else the reference's "last robot" quirk
// (Optimization3D_multi.h:730,792).
struct RobotLsArgs {
  RobotEnergyArgs e;
  int kte;            // trial points valid in this round: k0 .. kte-1 (stride of the tables stays e.KT)
  const double *wolfe, *ptime, *tdir;
  double *step, *ptrial, *tstep, *ttime;
  int* done;
  int wolfe_idx, slot;
  DevCounts* dc;
};

__global__ void __launch_bounds__(32 * TOB_LS_TRIALS) k_robot_ls(RobotLsArgs b) {
  const RobotEnergyArgs& a = b.e;
  const int robot = a.robot_begin + blockIdx.x;
  if (b.done[robot]) return;
  const int lane = threadIdx.x & 31, k = a.k0 + (threadIdx.x >> 5);   // one warp per trial point
  if (k < b.kte) {
    double e = 0;
    const bool bad = a.bad[robot * a.KT + k] != 0;            // the row sums of an infeasible trial are incomplete: unused
    // four rows of the lane per step, every load of the four issued before the first use (the loop was a chain of dependent
    // L2 round trips: plane count -> partials, 16 times over for a 512-row trajectory); summed in the same order as before
    for (int tr0 = lane; tr0 < a.n_tr && !bad; tr0 += 128) {
      int V[4];
      double p0[4], pb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int tr = tr0 + 32 * i;
        V[i] = 0; p0[i] = 0; pb[i] = 0;
        if (tr < a.n_tr) {
          const int row = robot * a.n_tr + tr;
          const size_t o = ((size_t)k * a.rows_all + row) * EN_REC;
          V[i] = en_vwarps(a.pl_off[row + 1] - a.pl_off[row]);
          p0[i] = a.row_e[o];
          pb[i] = a.row_e[o + EN_VMAX];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int tr = tr0 + 32 * i;
        if (tr < a.n_tr) {
          double sum = p0[i];
          if (V[i] > 1) {
            const double* p = a.row_e + ((size_t)k * a.rows_all + (robot * a.n_tr + tr)) * EN_REC;
            for (int v = 1; v < V[i]; v++) sum += p[v];
          }
          e += a.lambda * sum + a.lambda * pb[i];
        }
      }
    }
    const double t = a.ttime[robot * a.KT + k];
    const double st = (a.dir && a.tstep) ? a.tstep[robot * a.KT + k] : 0.0;
    for (int sp = lane; sp < a.P; sp += 32) {
      const double* C = a.convert + (size_t)36 * sp;
      double acc = 0;
      for (int ax = 0; ax < 3; ax++) {
        double bz[6];
        for (int q = 0; q < 6; q++) {
          size_t g = (size_t)robot * 3 * a.T + (size_t)ax * a.T + 3 * sp + q;
          bz[q] = a.dir ? __dadd_rn(a.spline[g], __dmul_rn(st, a.dir[g])) : a.spline[g];
        }
        for (int r = 0; r < 6; r++) {
          double cx = 0;
          for (int q = 0; q < 6; q++) cx += C[r + 6 * q] * bz[q];
          size_t s = (size_t)robot * 18 * a.P + (size_t)ax * 6 * a.P + 6 * sp + r;
          double pd = cx - a.pslack[s];
          acc += a.mu / 2.0 * pd * pd + a.plambda[s] * pd;
        }
      }
      double dt = t - a.tslack[(size_t)robot * a.P + sp];
      acc += a.mu / 2.0 * dt * dt + a.tlambda[(size_t)robot * a.P + sp] * dt;
      e += acc;
    }
    e = warp_sum(e);
    if (lane == 0) a.e_out[robot * a.KT + k] = bad ? INFINITY : e;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int KT = a.KT, u = robot, kte = b.kte;
  const double w = b.wolfe[b.wolfe_idx < 0 ? u : b.wolfe_idx];
  const double e0 = a.e_out[u * KT];
  for (int k = 1; k < kte; k++) {
    const double s = b.tstep[u * KT + k];
    if (!(e0 - 1e-4 * w * s < a.e_out[u * KT + k])) {
      // accepted rung of the ladder that started at step[u] (work counter: the line-search policy is tuned on this histogram)
      const int rung = (int)lrint(log(s / b.step[u]) / log(0.8));
      atomicAdd(&b.dc->ls_hist[rung < 0 ? 0 : (rung > 7 ? 7 : rung)], 1ull);
      b.step[u] = s;
      b.ptrial[u] = b.ttime[u * KT + k];
      b.done[u] = 1;
      return;
    }
  }
  double s = b.tstep[u * KT + kte - 1] * 0.8;        // continue the ladder below the last rung that was evaluated
  for (int k = 1; k < KT; k++) a.bad[u * KT + k] = 0;   // the trial slots are laid out anew: clear their infeasibility flags
  for (int k = 1; k < KT; k++) {
    b.tstep[u * KT + k] = s;
    // unfused like k_ls_init (api.cu is built without FMA contraction, this file with): a rung must get the same trial time
    // whichever kernel lays it out, or the result would depend on the line-search policy in the last bit
    b.ttime[u * KT + k] = __dadd_rn(b.ptime[u], __dmul_rn(s, b.tdir[u]));
    s *= 0.8;
  }
  atomicAdd(&b.dc->ls_pending[b.slot], 1);
}

// extra CTAs of an energy launch (they share the listed virtual warps of the heavy rows; those without an item leave at
// once): few when the launch has few rows (latency regime), up to 64 per SM when it has many -- measured on the
// 1024-problem batch: with 4 per SM the 36 k listed virtual warps ran on 8 warps per SM
static int energy_extra(tob_ctx* c, int nrows) {
  if (const char* e = getenv("TRAJOPT_B200_EN_EXTRA")) { int v = atoi(e); if (v >= 1 && v <= 256) return v * c->sm_count; }
  const int lo = 4 * c->sm_count, hi = 64 * c->sm_count, want = nrows;     // measured: 8192 rows 0.86 ms at 32+ per SM, 1.37 at 4
  return want < lo ? lo : (want > hi ? hi : want);
}

// list the virtual warps v >= 1 of the current plane CSR (c->pl_off): after every plane pass / plane upload
// reset: the item counter is not already zero (the plane pass zeroes it in k_np_top)
int energy_items(tob_ctx* c, bool reset) {
  const int rows = c->rows_all();
  TOB_CUDA(c, c->en_items.ensure((size_t)(EN_VMAX - 1) * rows + 8));    // at most EN_VMAX - 1 listed virtual warps per row
  if (reset) TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->n_en_items, 0, sizeof(uint32_t), c->stream));
  TOB_CUDA(c, c->en_item_base.ensure((size_t)rows + 1));
  k_en_items<<<div_up(rows, 256), 256, 0, c->stream>>>(c->pl_off.p, rows, c->en_items.p, c->en_item_base.p, c->dc.p);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

static int energy_buffers(tob_ctx* c, int KT) {
  TOB_CUDA(c, c->row_e.ensure((size_t)EN_REC * c->rows_all() * KT));
  TOB_CUDA(c, c->row_bad.ensure((size_t)c->n_robots() * TOB_LS_TRIALS + 1));
  return 0;
}

// partial plane sums of the listed parts of heavy rows (k_row_grad): one 54-vector per item
static int grad_buffers(tob_ctx* c) {
  TOB_CUDA(c, c->en_items.ensure((size_t)(EN_VMAX - 1) * c->rows_all() + 8));
  TOB_CUDA(c, c->en_item_base.ensure((size_t)c->rows_all() + 1));
  TOB_CUDA(c, c->gpart.ensure((size_t)54 * ((size_t)(EN_VMAX - 1) * c->rows_all() + 8)));
  return 0;
}

// Energies of trial points k0..k1-1 of robots [rb,re): trial spline = s_spline + tstep[u*KT+k]*dir, time ttime[u*KT+k].
// dir/tstep may be null (current point).  e_dev: robots x KT.
int energy_trials(tob_ctx* c, int rb, int re, const double* dir, const double* tstep, const double* ttime, int KT, int k0,
                  int k1, double* e_dev) {
  const int rows_all = c->rows_all();
  TOB_TRY(energy_buffers(c, KT));
  TOB_CUDA(c, cudaMemsetAsync(c->row_bad.p + (size_t)rb * KT, 0, (size_t)(re - rb) * KT * sizeof(int), c->stream));
  EnergyArgs a;
  a.spline = c->s_spline.p; a.dir = dir; a.tstep = tstep; a.ttime = ttime; a.basis = c->d_basis.p;
  a.pl = c->pl.p; a.pl_off = c->pl_off.p; a.weight = c->d_weight.p;
  a.margin = c->prm.margin; a.vel_limit = c->prm.vel_limit; a.acc_limit = c->prm.acc_limit;
  a.n_tr = c->n_tr; a.res = c->prm.res; a.T = c->T; a.row_begin = rb * c->n_tr; a.rows_all = rows_all; a.KT = KT; a.k0 = k0;
  a.row_e = c->row_e.p; a.bad = c->row_bad.p; a.done = nullptr; a.dc = c->dc.p; a.items = c->en_items.p; a.pending = nullptr;
  const int nrows = (re - rb) * c->n_tr, nk = k1 - k0;
  a.n_rows = nrows;
  {
    Prof prof(c, K_ROW_ENERGY);
    a.nk = nk;
    if (nk > EN_MAXT) return fail_msg(c, "energy_trials: too many trial points in one launch");
    launch_row_energy(c, a, nrows + energy_extra(c, nrows), nk, nrows);
    TOB_LAUNCH_CHECK(c);
  }
  RobotEnergyArgs b;
  b.spline = c->s_spline.p; b.dir = dir; b.tstep = tstep; b.ttime = ttime;
  b.pslack = c->s_pslack.p; b.tslack = c->s_tslack.p; b.plambda = c->s_plambda.p; b.tlambda = c->s_tlambda.p;
  b.convert = c->d_convert.p; b.row_e = c->row_e.p; b.pl_off = c->pl_off.p; b.bad = c->row_bad.p;
  b.lambda = c->prm.lambda; b.mu = c->prm.mu; b.n_tr = c->n_tr; b.P = c->prm.piece_num; b.T = c->T; b.robot_begin = rb;
  b.rows_all = rows_all; b.KT = KT; b.k0 = k0; b.e_out = e_dev;
  {
    Prof prof(c, K_ROBOT_ENERGY);
    k_robot_energy<<<dim3(re - rb, nk), 128, 0, c->stream>>>(b);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

// buffers of a line search (the infeasibility flags of the trial slots are cleared by k_ls_init, and by k_robot_ls for the
// slots it lays out anew)
int line_search_begin(tob_ctx* c, int rb, int re) {
  (void)rb; (void)re;
  return energy_buffers(c, TOB_LS_TRIALS);
}

// k0e: first trial the ENERGY launch evaluates (k0, or k0 + 1 = 1 when slot 0 was already written by the gradient pass at
// the same point); the robot kernel always sums trials k0 .. kte-1
int line_search_round(tob_ctx* c, int rb, int re, int wolfe_idx, int k0, int kte, int slot, int k0e) {
  const int rows_all = c->rows_all(), KT = TOB_LS_TRIALS;
  if (kte < 2 || kte > KT || k0 >= kte || k0e < k0 || k0e >= kte) return fail_msg(c, "line_search_round: bad trial range");
  EnergyArgs a;
  a.spline = c->s_spline.p; a.dir = c->s_dir.p; a.tstep = c->s_tstep.p; a.ttime = c->s_ttime.p; a.basis = c->d_basis.p;
  a.pl = c->pl.p; a.pl_off = c->pl_off.p; a.weight = c->d_weight.p;
  a.margin = c->prm.margin; a.vel_limit = c->prm.vel_limit; a.acc_limit = c->prm.acc_limit;
  a.n_tr = c->n_tr; a.res = c->prm.res; a.T = c->T; a.row_begin = rb * c->n_tr; a.rows_all = rows_all; a.KT = KT; a.k0 = k0e;
  a.row_e = c->row_e.p; a.bad = c->row_bad.p; a.done = c->s_done.p; a.dc = c->dc.p; a.items = c->en_items.p;
  // rounds 1.. of the sequence launched ahead: slot r counts the robots that round r left backtracking (k_robot_ls)
  a.pending = (slot >= 1 && slot < TOB_LS_MAXROUNDS) ? &c->dc.p->ls_pending[slot - 1] : nullptr;
  const int nrows = (re - rb) * c->n_tr;
  a.n_rows = nrows;
  {
    Prof prof(c, K_ROW_ENERGY);
    a.nk = kte - k0e;
    launch_row_energy(c, a, nrows + energy_extra(c, nrows), kte - k0e, nrows);
    TOB_LAUNCH_CHECK(c);
  }
  RobotLsArgs b;
  b.e.spline = c->s_spline.p; b.e.dir = c->s_dir.p; b.e.tstep = c->s_tstep.p; b.e.ttime = c->s_ttime.p;
  b.e.pslack = c->s_pslack.p; b.e.tslack = c->s_tslack.p; b.e.plambda = c->s_plambda.p; b.e.tlambda = c->s_tlambda.p;
  b.e.convert = c->d_convert.p; b.e.row_e = c->row_e.p; b.e.pl_off = c->pl_off.p; b.e.bad = c->row_bad.p;
  b.e.lambda = c->prm.lambda; b.e.mu = c->prm.mu; b.e.n_tr = c->n_tr; b.e.P = c->prm.piece_num; b.e.T = c->T; b.e.robot_begin = rb;
  b.e.rows_all = rows_all; b.e.KT = KT; b.e.k0 = k0; b.e.e_out = c->s_etr.p;
  b.wolfe = c->s_wolfe.p; b.ptime = c->s_ptime.p; b.tdir = c->s_tdir.p;
  b.step = c->s_step.p; b.ptrial = c->s_ptrial.p; b.tstep = c->s_tstep.p; b.ttime = c->s_ttime.p; b.done = c->s_done.p;
  b.wolfe_idx = wolfe_idx; b.slot = slot; b.dc = c->dc.p; b.kte = kte;
  {
    Prof prof(c, K_ROBOT_ENERGY);
    k_robot_ls<<<re - rb, 32 * (kte - k0), 0, c->stream>>>(b);
    TOB_LAUNCH_CHECK(c);
  }
  c->ctr.line_search_trials += (uint64_t)(re - rb) * (kte - 1);
  return 0;
}

// ---- gradient: per-row terms ---------------------------------------------------------------------------------------
struct RowGradArgs {
  const double* P;
  const double* pl;
  const uint32_t* pl_off;
  const double* weight;
  const double* ptime;     // per robot (current)
  double margin, vel_limit, acc_limit;
  int n_tr, row_begin, n_rows;
  double* terms;           // rows x ROW_REC: bound terms, and the plane terms of part 0
  // rows with more than EN_VPLANES planes are cut into V = en_vwarps(planes) parts (chunks of 128 planes dealt round-robin)
  // like the energy: part 0 runs in the row's own CTA, parts 1 .. V-1 are listed items shared by the extra CTAs; the plane
  // sums of item i (54 doubles) go to gpart[i] and are added in part order by whoever reads the row (row_plane_term)
  const uint32_t* items;
  double* gpart;           // items x 54
  DevCounts* dc;           // barrier_terms counter, n_en_items
  // by-product: barrier energy of the CURRENT point (trial slot 0 of the line search that follows): the logarithms of the
  // gradient are the logarithms of the energy, so the line search does not evaluate its starting point again
  double* row_e;           // trials x rows_all x EN_REC (slot 0 is written: one partial per part), may be null
  int rows_all;
};

// part v of `row` by the whole CTA (128 threads): sums of the plane terms -> s_red[4][54] per warp, energy -> s_en[4].
// A plane is shared by TWO threads: lanes 0-15 of a warp take control points 0..2 of 16 planes, lanes 16-31 control points
// 3..5 of the same planes (the second load of a plane is an L1 hit).  27 accumulators per thread instead of 54: the kernel
// ran at 234 registers = 8 warps per SM with the FP64 pipe a third busy, waiting on its own dependent chains.  A chunk of 128
// planes takes the CTA two steps of 64; the sums stay a function of the row's plane count only.
__device__ __forceinline__ void grad_part(const RowGradArgs& a, int row, int v, int V, const double* sP, double (*s_red)[54],
                                          double* s_en, const LogTabEntry* s_lt) {
  const int tr = row % a.n_tr;
  const double w = a.weight[tr], m = a.margin, inv_m = 1.0 / m;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, half = lane >> 4;
  const uint32_t pslot = (uint32_t)wp * 16u + (uint32_t)(lane & 15);      // plane of this thread inside a step of 64
  double acc[27];
  double en = 0;
  unsigned n_act = 0;
#pragma unroll
  for (int i = 0; i < 27; i++) acc[i] = 0;
  double cp[9];                 // the thread's three control points: x, y, z
#pragma unroll
  for (int jj = 0; jj < 3; jj++) { cp[jj] = sP[3 * half + jj]; cp[3 + jj] = sP[3 * half + jj + 6]; cp[6 + jj] = sP[3 * half + jj + 12]; }
  const uint32_t k0 = a.pl_off[row], k1 = a.pl_off[row + 1];
  const uint32_t first = k0 + (uint32_t)v * 128u, stride = 128u * (uint32_t)V;
  // step i: planes first + (i / 2) * stride + (i % 2) * 64 + [0, 64)
  double4 nxt = make_double4(0, 0, 0, 0);
  if (first + pslot < k1) nxt = *reinterpret_cast<const double4*>(a.pl + (size_t)4 * (first + pslot));
  for (uint32_t cb = first, i = 0; cb < k1; i++, cb += (i & 1u) ? 0u : stride) {     // uniform in the CTA
    const uint32_t k = cb + (i & 1u) * 64u + pslot;
    const double4 pl = nxt;       // the next plane of this thread is in flight while this one is evaluated
    {
      const uint32_t cbn = cb + ((i & 1u) ? stride : 0u), kn = cbn + ((i + 1u) & 1u) * 64u + pslot;
      if (cbn < k1 && kn < k1) nxt = *reinterpret_cast<const double4*>(a.pl + (size_t)4 * kn);
    }
    if (k >= k1) continue;
    const double cxx = pl.x * pl.x, cxy = pl.x * pl.y, cxz = pl.x * pl.z, cyy = pl.y * pl.y, cyz = pl.y * pl.z, czz = pl.z * pl.z;
    // branch-free like k_row_energy: a term outside the band contributes e1 = e2 = 0 through log(1) and dm = 0, so the three
    // log / reciprocal chains of the thread are independent and interleave
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
      const double d = cp[jj] * pl.x + cp[3 + jj] * pl.y + cp[6 + jj] * pl.z + pl.w;
      const bool act = d < m;
      n_act += act;
      const double lg = tob_log_pos(act ? d * inv_m : 1.0, s_lt), dm = act ? d - m : 0.0, id = 1.0 / (act ? d : 1.0);
      const double e1 = -w * (2 * dm * lg + dm * dm * id);
      const double e2 = -w * (2 * lg + 4 * dm * id - dm * dm * id * id);
      en += (dm * dm) * lg;
      double* q = acc + 9 * jj;
      q[0] += e2 * cxx; q[1] += e2 * cxy; q[2] += e2 * cxz; q[3] += e2 * cyy; q[4] += e2 * cyz; q[5] += e2 * czz;
      q[6] += e1 * pl.x; q[7] += e1 * pl.y; q[8] += e1 * pl.z;
    }
  }
  for (int o = 16; o; o >>= 1) n_act += __shfl_xor_sync(0xffffffffu, n_act, o);
  if (lane == 0 && n_act) atomicAdd(&a.dc->barrier_terms, (unsigned long long)n_act);
  if (first + 16u * wp < k1) {         // this warp streamed at least one plane (warp-uniform)
#pragma unroll
    for (int i = 0; i < 27; i++) {
      double x = acc[i];
      for (int o = 8; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);       // inside the half warp
      if ((lane & 15) == 0) s_red[wp][27 * half + i] = x;                        // control point 3 * half + i / 9, value i % 9
    }
    en = warp_sum(en);
    if (lane == 0) s_en[wp] = en;
  } else {
    for (int i = lane; i < 54; i += 32) s_red[wp][i] = 0.0;
    if (lane == 0) s_en[wp] = 0.0;
  }
}

// grid = n_rows + extra CTAs.  CTA b < n_rows: part 0 of row b and its bound terms; the extra CTAs share the listed parts.
__global__ void __launch_bounds__(128, 4) k_row_grad(RowGradArgs a) {
  __shared__ double sP[18];
  __shared__ double s_red[4][54];
  __shared__ double s_gt[9], s_ht[9], s_eb[9], s_en[4];
  __shared__ LogTabEntry s_lt[TOB_LOGTAB_N];
  if (a.dc->overflow) return;      // the plane CSR of this iteration was not built (see k_row_energy)
  load_logtab(s_lt);               // visible after the first barrier below (both paths have one before grad_part)
  if ((int)blockIdx.x >= a.n_rows) {
    const uint32_t n_items = a.dc->n_en_items, extra = gridDim.x - (uint32_t)a.n_rows;
    for (uint32_t idx = blockIdx.x - (uint32_t)a.n_rows; idx < n_items; idx += extra) {
      const uint32_t it = a.items[idx];
      const int row = (int)(it >> 3), v = (int)(it & 7u) + 1;
      if (row < a.row_begin || row >= a.row_begin + a.n_rows) continue;     // uniform
      __syncthreads();             // sP / s_red of the previous item have been consumed
      if (threadIdx.x < 18) sP[threadIdx.x] = a.P[(size_t)18 * row + threadIdx.x];
      __syncthreads();
      const int V = en_vwarps(a.pl_off[row + 1] - a.pl_off[row]);
      grad_part(a, row, v, V, sP, s_red, s_en, s_lt);
      __syncthreads();
      if (threadIdx.x < 54) {
        const int i = threadIdx.x;
        a.gpart[(size_t)54 * idx + i] = (s_red[0][i] + s_red[1][i]) + (s_red[2][i] + s_red[3][i]);
      }
      if (threadIdx.x == 64 && a.row_e)
        a.row_e[(size_t)row * EN_REC + v] = -a.weight[row % a.n_tr] * ((s_en[0] + s_en[1]) + (s_en[2] + s_en[3]));
    }
    return;
  }
  const int row = a.row_begin + blockIdx.x;
  const int robot = row / a.n_tr, tr = row - robot * a.n_tr;
  if (threadIdx.x < 18) sP[threadIdx.x] = a.P[(size_t)18 * row + threadIdx.x];
  __syncthreads();
  const double w = a.weight[tr], m = a.margin;
  const int V = en_vwarps(a.pl_off[row + 1] - a.pl_off[row]);
  grad_part(a, row, 0, V, sP, s_red, s_en, s_lt);
  // bound terms, threads 0..8
  double* out = a.terms + (size_t)ROW_REC * row;
  if (threadIdx.x < 9) {
    const double t = a.ptime[robot];
    double px, py, pz, dn, d, g_t = 0, h_t = 0, coef, e3k, eb = 0;
    double M3[6] = {0, 0, 0, 0, 0, 0}, g3[3] = {0, 0, 0}, pg3[3] = {0, 0, 0};
    bool vel = threadIdx.x < 5;
    double val;   // v or a of the reference
    if (vel) {
      int j = threadIdx.x;
      px = sP[j + 1] - sP[j]; py = sP[j + 7] - sP[j + 6]; pz = sP[j + 13] - sP[j + 12];
      dn = sqrt(px * px + py * py + pz * pz);
      val = 5 * dn / w;
      d = a.vel_limit - val / t;
      coef = -5 / (w * t);
    } else {
      int j = threadIdx.x - 5;
      px = sP[j + 2] - 2 * sP[j + 1] + sP[j]; py = sP[j + 8] - 2 * sP[j + 7] + sP[j + 6]; pz = sP[j + 14] - 2 * sP[j + 13] + sP[j + 12];
      dn = sqrt(px * px + py * py + pz * pz);
      val = 20 * dn / (w * w);
      d = a.acc_limit - val / (t * t);
      coef = -20 / ((w * t) * (w * t));
    }
    if (d < m) {
      double lg = log(d / m), dm = d - m;
      double e1 = -w * (2 * dm * lg + dm * dm / d);
      double e2 = -w * (2 * lg + 4 * dm / d - dm * dm / (d * d));
      eb = -w * dm * dm * lg;
      if (vel) {
        g_t = e1 * val / (t * t);
        h_t = -2 * e1 * val / (t * t * t) + e2 * val * val / (t * t * t * t);
        e3k = -e1 / t + e2 * (a.vel_limit - d) / t;
      } else {
        g_t = 2 * e1 * val / (t * t * t);
        h_t = -6 * e1 * val / (t * t * t * t) + 4 * e2 * val * val / (t * t * t * t * t * t);
        e3k = -2 * e1 / t + 2 * e2 * (a.acc_limit - d) / t;
      }
      double dp[3] = {coef * px / dn, coef * py / dn, coef * pz / dn};
      double i1 = 1.0 / dn, i3 = 1.0 / (dn * dn * dn);
      double pv[3] = {px, py, pz};
      // h_p = coef * (I/dn - p p^T / dn^3)
      int q = 0;
      for (int r = 0; r < 3; r++)
        for (int s = r; s < 3; s++) {
          double hp = coef * ((r == s ? i1 : 0.0) - pv[r] * pv[s] * i3);
          M3[q++] = e2 * dp[r] * dp[s] + e1 * hp;
        }
      for (int r = 0; r < 3; r++) { g3[r] = e1 * dp[r]; pg3[r] = e3k * dp[r]; }
    }
    double* o = out + (size_t)TERM_SZ * (6 + threadIdx.x);
    for (int i = 0; i < 6; i++) o[i] = M3[i];
    for (int i = 0; i < 3; i++) { o[6 + i] = g3[i]; o[9 + i] = pg3[i]; }
    s_gt[threadIdx.x] = g_t; s_ht[threadIdx.x] = h_t; s_eb[threadIdx.x] = eb;
  }
  __syncthreads();
  if (threadIdx.x < 54) {
    int i = threadIdx.x;
    double x = (s_red[0][i] + s_red[1][i]) + (s_red[2][i] + s_red[3][i]);
    int j = i / 9, q = i - 9 * j;
    double* o = out + (size_t)TERM_SZ * j;
    o[q] = x;
    if (q < 3) o[9 + q] = 0.0;   // plane terms carry no time coupling
  }
  if (threadIdx.x == 64) {
    double g = 0, h = 0, eb = 0;
    for (int i = 0; i < 9; i++) { g += s_gt[i]; h += s_ht[i]; eb += s_eb[i]; }
    out[ROW_TERMS * TERM_SZ] = g;
    out[ROW_TERMS * TERM_SZ + 1] = h;
    if (a.row_e) {                  // record of (trial 0, row): partial 0 of the plane energy (parts >= 1: the extra CTAs)
      double* re = a.row_e + (size_t)row * EN_REC;
      re[0] = -w * ((s_en[0] + s_en[1]) + (s_en[2] + s_en[3]));
      re[EN_VMAX] = eb;
    }
  }
}

// plane-term value q (< 9) of term tt (< 6) of `row`: part 0 from the row record, parts 1 .. V-1 from the item records
__device__ __forceinline__ double row_plane_term(const double* __restrict__ terms, const double* __restrict__ gpart,
                                                 const uint32_t* __restrict__ item_base, const uint32_t* __restrict__ pl_off, size_t row,
                                                 int tt, int q) {
  double x = terms[(size_t)ROW_REC * row + TERM_SZ * tt + q];
  const int V = en_vwarps(pl_off[row + 1] - pl_off[row]);
  if (V > 1) {
    const size_t base = item_base[row];
    for (int v = 1; v < V; v++) x += gpart[(size_t)54 * (base + v - 1) + 9 * tt + q];
  }
  return x;
}

// ---- block-cooperative PSD projection of one 19x19 block (Gradient_admm.h:40-53) ------------------------------------------
// The reference tests the block with Eigen's LLT and, when that fails, shifts it by (-lambda_min + 0.01) I if lambda_min < 0.
// Away from the boundary the LLT outcome IS the sign of lambda_min (a pivot <= 0 needs lambda_min <~ n eps |H|), so the
// order is turned around: lambda_min first, the Cholesky test only when |lambda_min| is within 1e-11 of the scale of H.
// Once the bound barriers are active most blocks of an iteration are indefinite, and "Cholesky test, then eigenvalues" cost
// 52 us per launch against 21 us when every block was SPD.
//   lambda_min      warp 0: Householder tridiagonalisation with lane j keeping column j of the symmetric block in registers
//                   (row k of the matrix is spread over the lanes by symmetry, A v needs no reduction, the two rank-1 vectors
//                   go through shared memory with one warp barrier each), then Sturm-count multisection by the whole CTA:
//                   384 shifts per round on the division-free recurrence
//   Cholesky test   (borderline blocks only) thread per element, one barrier per pivot; same operands as Eigen's unblocked LLT
// Returns 0 = SPD, 1 = shifted by (-lambda_min + 0.01) I in place, 2 = LLT failed but lambda_min >= 0.  All 384 threads call.
__device__ __forceinline__ void warp_tridiag19(const double* s_H, double* s_d, double* s_e, double* s_v, double* s_w) {
  constexpr int N = 19;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  double a[N];
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = lane < N ? s_H[i + N * lane] : 0.0;
#pragma unroll
  for (int k = 0; k + 2 < N; k++) {
    // x = A(k, k+1..): lane j > k holds x_j = a[k] (symmetry)
    const double xj = (lane > k && lane < N) ? a[k] : 0.0;
    double sigma = xj * xj;
#pragma unroll
    for (int o = 16; o; o >>= 1) sigma += __shfl_xor_sync(full, sigma, o);
    const double x0 = __shfl_sync(full, a[k], k + 1);
    const double tail = sigma - x0 * x0;
    if (lane == k) s_d[k] = a[k];
    if (!(tail > 0)) {                                     // column already tridiagonal (uniform)
      if (lane == k) s_e[k] = x0;
      continue;
    }
    const double alpha = (x0 >= 0 ? -1.0 : 1.0) * sqrt(sigma);
    const double vk1 = x0 - alpha;
    const double beta = 2.0 / (tail + vk1 * vk1);
    const double vj = lane == k + 1 ? vk1 : xj;
    __syncwarp();
    if (lane < N) s_v[lane] = vj;
    __syncwarp();
    // p = beta A v: lane j owns column j = row j
    double p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
    for (int i = k + 1; i < N; i += 3) {
      p0 += a[i] * s_v[i];
      if (i + 1 < N) p1 += a[i + 1] * s_v[i + 1];
      if (i + 2 < N) p2 += a[i + 2] * s_v[i + 2];
    }
    const double pj = (lane > k && lane < N) ? beta * ((p0 + p1) + p2) : 0.0;
    double kk = pj * vj;
#pragma unroll
    for (int o = 16; o; o >>= 1) kk += __shfl_xor_sync(full, kk, o);
    kk *= 0.5 * beta;
    const double wj = pj - kk * vj;
    if (lane < N) s_w[lane] = wj;
    __syncwarp();
    if (lane > k && lane < N) {
#pragma unroll
      for (int i = k + 1; i < N; i++) a[i] -= s_v[i] * wj + s_w[i] * vj;
    }
    if (lane == k) s_e[k] = alpha;
  }
  if (lane == N - 2) { s_d[N - 2] = a[N - 2]; s_e[N - 2] = a[N - 1]; }
  if (lane == N - 1) { s_d[N - 1] = a[N - 1]; s_e[N - 1] = 0.0; }
}

// NWP = warps of the CTA (12, or 4 where many blocks are in flight: the multisection then evaluates three of the 384 shifts of a
// round per thread and the Cholesky test three elements per thread -- the same shifts and the same operations, so the result
// does not depend on the variant)
template <int NWP>
__device__ __forceinline__ int cta_psd_shift19(double* s_H) {
  constexpr int N = 19, NW = 12, SP = NW / NWP;
  __shared__ double s_col[2][N + 1], s_v[N + 1], s_w[N + 1], s_d[N + 1], s_e[N + 1], s_e2[N + 1];
  __shared__ int s_first[2][NW];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned full = 0xffffffffu;
  if (w == 0) warp_tridiag19(s_H, s_d, s_e, s_v, s_w);
  __syncthreads();
  // Gershgorin bounds of the tridiagonal matrix; lambda_min <= min d_i
  double lo = INFINITY, hi = INFINITY, sc = 0.0;
  if (lane < N) {
    const double r = fabs(s_e[lane]) + (lane > 0 ? fabs(s_e[lane - 1]) : 0.0);
    lo = s_d[lane] - r;
    hi = s_d[lane];
    sc = fabs(s_d[lane]) + r;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(full, lo, o));
    hi = fmin(hi, __shfl_xor_sync(full, hi, o));
    sc = fmax(sc, __shfl_xor_sync(full, sc, o));
  }
  double mn = hi;
  if (hi > lo) {
    // Sturm counts on the division-free recurrence q_i = (d_i - x) q_{i-1} - e_{i-1}^2 q_{i-2}.  FP64 results take ~40 cycles
    // on this part, so the loop is laid out for the shortest dependent chain: d_i and e_i^2 come from shared memory at
    // static addresses (loads hoisted by the unrolling), (d_i - x) and
    // e^2 q_{i-2} do not depend on q_{i-1}, which leaves ONE fused multiply-add per step on the critical path; the sign
    // bookkeeping runs beside it and the range guard only every sixth step (19 factors of <= 1e12 cannot overflow between
    // two guards).  Rounds stop as soon as the bracket is below 1e-14 of the scale of the block.
    constexpr int NT = 32 * NW;
    if (tid < N) s_e2[tid] = tid ? s_e[tid - 1] * s_e[tid - 1] : 0.0;
    __syncthreads();
    const double* dd = s_d;          // static addresses in the unrolled loop: the loads are issued ahead of the chain
    const double* e2 = s_e2;
    const double stop = 1e-14 * sc;
    for (int round = 0; round < 8; round++) {
      const int b = round & 1;
#pragma unroll
      for (int sp = 0; sp < SP; sp++) {
        const int idx = sp * 32 * NWP + tid;               // shift of this thread: virtual warp sp * NWP + w of the 12
        const double x = lo + (hi - lo) * ((idx + 1) / (double)(NT + 1));
        int cnt = 0;
        double q0 = 1.0, q1 = dd[0] - x;
        bool neg = q1 < 0;                                   // sign of the last non-zero term
        if (neg) cnt++;
#pragma unroll
        for (int i = 1; i < N; i++) {
          const double t = e2[i] * q0;
          double q2 = fma(dd[i] - x, q1, -t);
          if (i % 6 == 0) {
            const double m = fabs(q2);
            if (m > 1e100) { q2 *= 1e-100; q1 *= 1e-100; }
            else if (m < 1e-100 && m > 0) { q2 *= 1e100; q1 *= 1e100; }
          }
          if (q2 != 0) {
            const bool n2 = q2 < 0;
            cnt += (n2 != neg);
            neg = n2;
          }
          q0 = q1; q1 = q2;
        }
        const unsigned mk = __ballot_sync(full, cnt >= 1);
        if (lane == 0) s_first[b][sp * NWP + w] = mk ? (sp * NWP + w) * 32 + (__ffs(mk) - 1) : -1;
      }
      __syncthreads();
      int f = -1;
      for (int ww = 0; ww < NW; ww++)
        if (s_first[b][ww] >= 0) { f = s_first[b][ww]; break; }
      const double width = hi - lo, base = lo;
      if (f < 0) lo = base + width * (NT / (double)(NT + 1));
      else {
        hi = base + width * ((f + 1) / (double)(NT + 1));
        if (f > 0) lo = base + width * (f / (double)(NT + 1));
      }
      if (!(hi - lo > stop)) break;
    }
    mn = 0.5 * (lo + hi);
  }
  const double band = 1e-11 * sc;
  bool llt_ok = mn > 0;
  if (fabs(mn) <= band) {
    // borderline: the reference's decision is the outcome of Eigen's unblocked LLT on these very operands
    constexpr int EPT = (N * N + 32 * NWP - 1) / (32 * NWP);     // elements per thread
    double aij[EPT];
#pragma unroll
    for (int q = 0; q < EPT; q++) { const int e = tid + q * 32 * NWP; aij[q] = e < N * N ? s_H[e] : 0.0; }
    llt_ok = true;
    __syncthreads();
    for (int k = 0; k < N; k++) {
      double* col = s_col[k & 1];
#pragma unroll
      for (int q = 0; q < EPT; q++) {
        const int e = tid + q * 32 * NWP, i = e % N, j = e / N;
        if (e < N * N && j == k && i >= k) col[i] = aij[q];
      }
      __syncthreads();
      const double x = col[k];
      if (!(x > 0)) { llt_ok = false; break; }             // uniform
      const double lkk = sqrt(x);
#pragma unroll
      for (int q = 0; q < EPT; q++) {
        const int e = tid + q * 32 * NWP, i = e % N, j = e / N;
        if (e < N * N && j > k && i >= j) aij[q] -= (col[i] / lkk) * (col[j] / lkk);
      }
    }
  }
  if (llt_ok) return 0;
  if (mn < 0) {
    __syncthreads();
    if (tid < N) s_H[tid * (N + 1)] = s_H[tid * (N + 1)] - mn * 1.0 + 0.01 * 1.0;
    return 1;
  }
  return 2;
}

// ---- gradient: per-piece 19x19 block -----------------------------------------------------------------------------
struct PieceArgs {
  const double *terms, *basis, *convert;
  const double* gpart;            // partial plane sums of the heavy rows (k_row_grad)
  const uint32_t *item_base, *pl_off;
  const double *spline, *ptime, *pslack, *tslack, *plambda, *tlambda;
  double lambda, mu;
  int n_tr, res, P, T, robot_begin, project_psd;
  double *pc_g, *pc_h;   // (robots*P) x 19 / x 361 (col-major)
  int* pc_flag;          // 0 SPD, 1 shifted, 2 LLT failed but lambda_min >= 0
  DevCounts* dc;         // planes x gradient passes counter
};

template <int NWP>
__global__ void __launch_bounds__(32 * NWP) k_piece(PieceArgs a) {
  extern __shared__ double sm[];
  const int robot = a.robot_begin + blockIdx.x / a.P, sp = blockIdx.x % a.P;
  const int nterm = a.res * ROW_TERMS;
  double* s_a = sm;                       // nterm x 6   a-vectors
  double* s_t = s_a + nterm * 6;          // nterm x 12  term payloads
  double* s_H = s_t + nterm * TERM_SZ;    // 361
  double* s_L = s_H + 361;                // 361 scratch
  double* s_x = s_L + 361;                // 36: x1 (18) x2 (18) as [m][k]
  double* s_sc = s_x + 36;                // gt, ht
  const double* C = a.convert + (size_t)36 * sp;
  // the plane CSR of this iteration was not built (candidate overflow: the host grows the buffers and repeats the
  // iteration): pl_off / the listed parts are stale, nothing below may be indexed with them
  if (a.dc->overflow & TOB_OVF_RETRY) return;
  for (int i = threadIdx.x; i < nterm * 6; i += blockDim.x) {
    int term = i / 6, mm = i - 6 * term;
    int rr = term / ROW_TERMS, tt = term - ROW_TERMS * rr;
    const double* B = a.basis + (size_t)36 * (sp * a.res + rr);
    double v;
    if (tt < 6) v = B[tt + 6 * mm];
    else if (tt < 11) { int j = tt - 6; v = B[j + 1 + 6 * mm] - B[j + 6 * mm]; }
    else { int j = tt - 11; v = B[j + 2 + 6 * mm] - 2 * B[j + 1 + 6 * mm] + B[j + 6 * mm]; }
    s_a[i] = v;
  }
  for (int i = threadIdx.x; i < nterm * TERM_SZ; i += blockDim.x) {
    int term = i / TERM_SZ, q = i - TERM_SZ * term;
    int rr = term / ROW_TERMS, tt = term - ROW_TERMS * rr;
    size_t row = (size_t)robot * a.n_tr + sp * a.res + rr;
    s_t[i] = (tt < 6 && q < 9) ? row_plane_term(a.terms, a.gpart, a.item_base, a.pl_off, row, tt, q)
                               : a.terms[(size_t)ROW_REC * row + TERM_SZ * tt + q];
  }
  if (threadIdx.x < 36) {
    // x1 = C^T (C bz - p_slack), x2 = C^T lambda ; index [m][k]
    int which = threadIdx.x / 18, mk = threadIdx.x % 18, mm = mk / 3, k = mk % 3;
    double acc = 0;
    for (int r = 0; r < 6; r++) {
      size_t s = (size_t)robot * 18 * a.P + (size_t)k * 6 * a.P + 6 * sp + r;
      double val;
      if (which == 0) {
        double cx = 0;
        for (int kk = 0; kk < 6; kk++) cx += C[r + 6 * kk] * a.spline[(size_t)robot * 3 * a.T + (size_t)k * a.T + 3 * sp + kk];
        val = cx - a.pslack[s];
      } else val = a.plambda[s];
      acc += C[r + 6 * mm] * val;
    }
    s_x[threadIdx.x] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 37) a.dc->energy_plane_evals += a.dc->n_planes;
  if (threadIdx.x == 36) {
    double g = 0, h = 0;
    for (int rr = 0; rr < a.res; rr++) {
      size_t row = (size_t)robot * a.n_tr + sp * a.res + rr;
      g += a.terms[(size_t)ROW_REC * row + ROW_TERMS * TERM_SZ];
      h += a.terms[(size_t)ROW_REC * row + ROW_TERMS * TERM_SZ + 1];
    }
    s_sc[0] = g; s_sc[1] = h;
  }
  __syncthreads();
  const size_t pb = (size_t)robot * a.P + sp;
  double* G = a.pc_g + 19 * pb;
  // Hessian entries
  for (int e = threadIdx.x; e < 361; e += blockDim.x) {
    int r = e % 19, cc = e / 19;
    double v = 0;
    if (r < 18 && cc < 18) {
      int m1 = r / 3, k1 = r % 3, m2 = cc / 3, k2 = cc % 3;
      int lo = k1 < k2 ? k1 : k2, hi = k1 < k2 ? k2 : k1;
      int q = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);   // xx xy xz yy yz zz
      for (int t = 0; t < nterm; t++) v += s_a[6 * t + m1] * s_a[6 * t + m2] * s_t[TERM_SZ * t + q];
      v *= a.lambda;
      if (k1 == k2) {
        double ctc = 0;
        for (int rr = 0; rr < 6; rr++) ctc += C[rr + 6 * m1] * C[rr + 6 * m2];
        v += a.mu * ctc;
      }
    } else if (r == 18 && cc == 18) {
      v = a.lambda * s_sc[1] + a.mu;
    } else {
      int idx = r == 18 ? cc : r;
      int m1 = idx / 3, k1 = idx % 3;
      for (int t = 0; t < nterm; t++) v += s_a[6 * t + m1] * s_t[TERM_SZ * t + 9 + k1];
      v *= a.lambda;
    }
    s_H[e] = v;
  }
  if (threadIdx.x < 19) {
    int r = threadIdx.x;
    double v;
    if (r < 18) {
      int m1 = r / 3, k1 = r % 3;
      v = 0;
      for (int t = 0; t < nterm; t++) v += s_a[6 * t + m1] * s_t[TERM_SZ * t + 6 + k1];
      v = a.lambda * v + a.mu * s_x[r] + s_x[18 + r];
    } else {
      v = a.lambda * s_sc[0] + a.mu * (a.ptime[robot] - a.tslack[pb]) + a.tlambda[pb];
    }
    G[r] = v;
  }
  __syncthreads();
  // PSD projection of Gradient_admm.h:40-53 by the whole CTA
  if (a.project_psd) {
    const int flag = cta_psd_shift19<NWP>(s_H);
    if (threadIdx.x == 0) a.pc_flag[pb] = flag;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 361; e += blockDim.x) a.pc_h[361 * pb + e] = s_H[e];
}

// ---- function-level: one row's block in piece coordinates ------------------------------------------------------------
// which = 0: Gradient_admm::local_plane_barrier_gradient :331-407 (terms 0..5); which = 1: local_bound_gradient :409-572
// (terms 6..14, plus g_t, h_t and the mixed time column).  out: g[18] | H[324 col-major] | g_t | h_t | partgrad[18]
__global__ void k_row_expand(const double* __restrict__ terms_all, const double* __restrict__ gpart, const uint32_t* __restrict__ item_base,
                             const uint32_t* __restrict__ pl_off, int row, const double* __restrict__ B, int which,
                             double* __restrict__ out) {
  __shared__ double s_a[ROW_TERMS * 6];
  __shared__ double terms[ROW_REC];
  for (int i = threadIdx.x; i < ROW_REC; i += blockDim.x) {
    const int tt = i / TERM_SZ, q = i - TERM_SZ * tt;
    terms[i] = (tt < 6 && q < 9) ? row_plane_term(terms_all, gpart, item_base, pl_off, (size_t)row, tt, q) : terms_all[(size_t)ROW_REC * row + i];
  }
  const int t0 = which ? 6 : 0, t1 = which ? ROW_TERMS : 6;
  for (int i = threadIdx.x; i < ROW_TERMS * 6; i += blockDim.x) {
    const int tt = i / 6, mm = i - 6 * tt;
    double v;
    if (tt < 6) v = B[tt + 6 * mm];
    else if (tt < 11) { int j = tt - 6; v = B[j + 1 + 6 * mm] - B[j + 6 * mm]; }
    else { int j = tt - 11; v = B[j + 2 + 6 * mm] - 2 * B[j + 1 + 6 * mm] + B[j + 6 * mm]; }
    s_a[i] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 324; e += blockDim.x) {
    const int r = e % 18, cc = e / 18, m1 = r / 3, k1 = r % 3, m2 = cc / 3, k2 = cc % 3;
    const int lo = k1 < k2 ? k1 : k2, hi = k1 < k2 ? k2 : k1;
    const int q = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
    double v = 0;
    for (int t = t0; t < t1; t++) v += s_a[6 * t + m1] * s_a[6 * t + m2] * terms[TERM_SZ * t + q];
    out[18 + e] = v;
  }
  if (threadIdx.x < 18) {
    const int m1 = threadIdx.x / 3, k1 = threadIdx.x % 3;
    double g = 0, pg = 0;
    for (int t = t0; t < t1; t++) { g += s_a[6 * t + m1] * terms[TERM_SZ * t + 6 + k1]; pg += s_a[6 * t + m1] * terms[TERM_SZ * t + 9 + k1]; }
    out[threadIdx.x] = g;
    out[18 + 324 + 2 + threadIdx.x] = pg;
  }
  if (threadIdx.x == 32) {
    out[18 + 324] = which ? terms[ROW_TERMS * TERM_SZ] : 0.0;
    out[18 + 324 + 1] = which ? terms[ROW_TERMS * TERM_SZ + 1] : 0.0;
  }
}

// geo.P of robot 0 and the resident planes must be current; fills out_dev[362]
int row_blocks(tob_ctx* c, int tr, int which, double* out_dev) {
  TOB_CUDA(c, c->row_terms.ensure((size_t)ROW_REC * c->rows_all()));
  TOB_TRY(grad_buffers(c));
  RowGradArgs a;
  a.P = c->geo.P.p; a.pl = c->pl.p; a.pl_off = c->pl_off.p; a.weight = c->d_weight.p; a.ptime = c->s_ptime.p;
  a.margin = c->prm.margin; a.vel_limit = c->prm.vel_limit; a.acc_limit = c->prm.acc_limit;
  a.n_tr = c->n_tr; a.row_begin = tr; a.n_rows = 1; a.terms = c->row_terms.p; a.dc = c->dc.p; a.row_e = nullptr; a.rows_all = c->rows_all();
  a.items = c->en_items.p; a.gpart = c->gpart.p;
  k_row_grad<<<1 + energy_extra(c, 1), 128, 0, c->stream>>>(a);
  TOB_LAUNCH_CHECK(c);
  k_row_expand<<<1, 128, 0, c->stream>>>(c->row_terms.p, c->gpart.p, c->en_item_base.p, c->pl_off.p, tr, c->d_basis.p + (size_t)36 * tr, which,
                                         out_dev);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// geo.P must hold the CURRENT rows of robots [rb,re) (compute_rows without trial); planes resident.
int gradient_blocks(tob_ctx* c, int rb, int re, int project_psd) {
  int rows_total = c->n_robots() * c->n_tr;
  int P = c->prm.piece_num;
  TOB_CUDA(c, c->row_terms.ensure((size_t)ROW_REC * rows_total));
  TOB_CUDA(c, c->pc_g.ensure((size_t)19 * c->n_robots() * P));
  TOB_CUDA(c, c->pc_h.ensure((size_t)361 * c->n_robots() * P));
  TOB_CUDA(c, c->pc_flag.ensure((size_t)c->n_robots() * P));
  RowGradArgs a;
  a.P = c->geo.P.p; a.pl = c->pl.p; a.pl_off = c->pl_off.p; a.weight = c->d_weight.p; a.ptime = c->s_ptime.p;
  a.margin = c->prm.margin; a.vel_limit = c->prm.vel_limit; a.acc_limit = c->prm.acc_limit;
  a.n_tr = c->n_tr; a.row_begin = rb * c->n_tr; a.n_rows = (re - rb) * c->n_tr; a.terms = c->row_terms.p; a.dc = c->dc.p;
  TOB_TRY(energy_buffers(c, TOB_LS_TRIALS));
  TOB_TRY(grad_buffers(c));
  a.row_e = c->row_e.p; a.rows_all = rows_total; a.items = c->en_items.p; a.gpart = c->gpart.p;
  {
    Prof prof(c, K_ROW_GRAD);
    k_row_grad<<<a.n_rows + energy_extra(c, a.n_rows), 128, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  PieceArgs b;
  b.terms = c->row_terms.p; b.basis = c->d_basis.p; b.convert = c->d_convert.p;
  b.gpart = c->gpart.p; b.item_base = c->en_item_base.p; b.pl_off = c->pl_off.p;
  b.spline = c->s_spline.p; b.ptime = c->s_ptime.p; b.pslack = c->s_pslack.p; b.tslack = c->s_tslack.p;
  b.plambda = c->s_plambda.p; b.tlambda = c->s_tlambda.p;
  b.lambda = c->prm.lambda; b.mu = c->prm.mu; b.n_tr = c->n_tr; b.res = c->prm.res; b.P = P; b.T = c->T;
  b.robot_begin = rb; b.project_psd = project_psd;
  b.pc_g = c->pc_g.p; b.pc_h = c->pc_h.p; b.pc_flag = c->pc_flag.p; b.dc = c->dc.p;
  size_t smem = ((size_t)c->prm.res * ROW_TERMS * (6 + TERM_SZ) + 361 * 2 + 36 + 2) * sizeof(double);
  {
    Prof prof(c, K_PIECE);
    // many blocks: 128-thread CTAs (the tridiagonalisation of a block runs on ONE warp: with 384 threads and two CTAs per SM
    // two warps per SM worked while 22 waited at the barrier behind it; same results, see cta_psd_shift19)
    const char* e = getenv("TRAJOPT_B200_PIECE_CTA");
    const int small = e ? atoi(e) == 128 : (re - rb) * P >= 1024;
    if (small) k_piece<4><<<(re - rb) * P, 128, smem, c->stream>>>(b);
    else k_piece<12><<<(re - rb) * P, 384, smem, c->stream>>>(b);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

}  // namespace tob
