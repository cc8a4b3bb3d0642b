// narrow.cu -- per-pair FP64 narrowphase kernels (built with --fmad=false, see gjk.cuh):
//   * obstacle planes: 49-DOP filter -> GJK -> separating plane per (sub-segment, point) candidate
//     (Optimization3D_admm::separate_plane, Optimization3D_admm.h:69-197 with optimal_plane=0)
//   * inter-robot planes per time slot: box -> 49-DOP -> GJK -> 1-D Newton on d
//     (Optimization3D_multi::separate_self, Optimization3D_multi.h:237-342)
//   * CCD conservative step: swept 49-DOP gate + GJK on hull(P U P+sD) with the 0.8 ladder and a per-robot
//     atomicMax of the ladder exponent (Step::position_step, Step.h:21-110); inter-robot variants
//     (Step::self_step :184-256, Step::couple_self_step :112-182)
//   * packing of accepted planes into the per-row CSR the barrier kernels stream.
//   * persistent-plane mode ("optimal_plane": 1): sorted sparse set of live (row, point) planes, merged with the planes
//     of pairs that were not live yet and refined in place every iteration by Optimal_plane::optimal_cd
//     (Optimization3D_admm.h:126-193); inter-robot planes kept per (slot, pair) and refined by self_optimal_cd
//     (Optimization3D_multi.h:271-339).
#include <algorithm>
#include <utility>
#include <vector>

#include "ctx.cuh"
#include "bp.cuh"
#define TOB_GJK_INLINE
#include "gjk.cuh"
#include "optplane.cuh"

namespace tob {

#define TOB_MAX_LADDER TOB_LADDER

__device__ __forceinline__ void load_pts6(const double* __restrict__ src, double (*P)[3]) {
  // src: 6x3 column-major (18 doubles)
  for (int j = 0; j < 6; j++) { P[j][0] = src[j]; P[j][1] = src[j + 6]; P[j][2] = src[j + 12]; }
}

// ---- obstacle planes -------------------------------------------------------------------------------------------
// Candidates are processed in chunks of up to NP_CHUNK (one CTA iteration; a CTA takes its next chunk from an atomic
// ticket).  Inside a chunk: every thread runs the first axes of the 49-DOP gate for up to NP_PER candidates, the survivors
// are compacted (ballot + prefix), and the threads then run GJK + plane for the survivors, so the expensive divergent part
// executes on dense warps instead of on ~30 % of the lanes.  Per chunk the number of accepted planes goes to csum[chunk];
// k_np_top scans it and k_pack scatters.
struct NarrowArgs {
  DevCounts* dc;
  uint32_t cap;
  const uint32_t *cand_pt, *cand_row;
  const double *px, *py, *pz;
  const double *P, *klo, *khi, *kdop;
  const float* kf;   // rows x TOB_KF_ROW: single-precision filter of the gate (segments.cu), with its centres kc (rows x 4)
  const double* kc;
  double dist, offset;
  double gate_skip;  // |witness| <= gate_skip: the rest of the gate is implied (see k_narrow); dist * (1 - 1e-6), or -1: never
  int gate1;         // axes of the gate that run on every candidate (a multiple of 7)
  double* cpl;       // cap x 4
  uint32_t* cflag;   // cap
  uint32_t* csum;    // chunks + 1
  const unsigned long long* live_key;   // persistent-plane mode: sorted keys of the live planes (else nullptr)
  uint32_t np_grid;                     // CTAs of the narrowphase grid (np_per_of)
};

// index of the first key >= x in the sorted array k[0..n)
__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long* __restrict__ k, uint32_t n, unsigned long long x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (k[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

#define NP_THREADS 128
#define NP_PER 8
#define NP_CHUNK (NP_THREADS * NP_PER)   // most candidates per CTA iteration; their k-DOP survivors fill the GJK phase densely

// Candidates per thread and chunk (1, 2, 4 or 8), chosen on the device from the candidate count so that a small query (one
// UAV: ~1e5 candidates) still spreads over every CTA of the grid instead of serialising two GJK rounds on half of them.
// k_narrow, k_np_top, k_pack and k_live_compact must agree: all derive it from (n_cand, np_grid).
// grid: CTAs the chunks should cover; its top three bits, when set, cap the result (TRAJOPT_B200_NP_CHUNK: A/B measurements)
__device__ __forceinline__ uint32_t np_per_of(uint32_t n, uint32_t grid_and_cap) {
  const uint32_t grid = grid_and_cap & 0x1fffffffu, cap = grid_and_cap >> 29;
  uint32_t per = 1;
  if ((n + 8 * NP_THREADS - 1) / (8 * NP_THREADS) >= 4 * grid) per = 8;      // at least ~20 chunks per CTA of the grid
  else if ((n + 4 * NP_THREADS - 1) / (4 * NP_THREADS) >= grid) per = 4;
  else if ((n + 2 * NP_THREADS - 1) / (2 * NP_THREADS) >= grid) per = 2;
  return cap && per > cap ? cap : per;
}
static uint32_t np_grid_of(const tob_ctx* c) {
  uint32_t g = (uint32_t)c->sm_count * 4;
  if (const char* e = getenv("TRAJOPT_B200_NP_CHUNK")) {      // 128 / 256 / 512: most candidates per chunk
    const int v = atoi(e);
    if (v == 128) g |= 1u << 29; else if (v == 256) g |= 2u << 29; else if (v == 512) g |= 4u << 29;
  }
  return g;
}

// block-wide stable compaction step: the threads with `keep` append `value` to list[count ...] in thread order; returns the
// new count (uniform).  Two barriers; s_w is scratch of NP_THREADS / 32 words.
__device__ __forceinline__ uint32_t np_append(bool keep, uint32_t value, uint32_t* list, uint32_t count, uint32_t* s_w) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t bm = __ballot_sync(0xffffffffu, keep);
  __syncthreads();                 // s_w of the previous step has been consumed
  if (lane == 0) s_w[w] = __popc(bm);
  __syncthreads();
  uint32_t base = count, tot = 0;
#pragma unroll
  for (int k = 0; k < NP_THREADS / 32; k++) {
    if (k < (int)w) base += s_w[k];
    tot += s_w[k];
  }
  if (keep) list[base + __popc(bm & ((1u << lane) - 1u))] = value;
  return count + tot;
}

#define NP_GATE1 14
// A pair gets a plane iff the 49-DOP gate passes (CCD::KDOPDCD) AND the GJK distance is <= dist (Separate::opengjk), both with
// the same gap dist = offset + margin (Optimization3D_admm.h:113-160).  The two tests are not independent: the gate is a
// conservative filter of the distance test.  The axes are unit vectors and the GJK witness is a convex combination of
// vertices of hull - point, so |witness| >= the true distance, and a true distance <= dist puts the point inside every slab
// widened by dist.  A pair with |witness| <= dist * (1 - 1e-6) therefore passes the gate by a margin (2e-7 for dist = 0.2)
// that is seven orders of magnitude above the rounding of either computation (1e-14), and the gate need not be evaluated for
// it; a pair with |witness| > dist gets no plane whatever the gate says.  Only inside the band in between -- and for a NaN
// witness -- is the gate's own arithmetic decisive, and there it is evaluated as the reference does.  So: the first NP_GATE1
// axes run on every candidate (they reject 40 % of them after one or two groups of 7 axes, far cheaper than GJK), the survivors
// are compacted and run GJK on dense warps, and the remaining axes run for the pairs in the band only (counted:
// DevCounts::np_band).  Before, 56 % of the candidates paid all 49 axes and then GJK.
// gate_skip = dist * (1 - 1e-6); -1 (TRAJOPT_B200_NP_BAND=0, tests) evaluates the remaining axes for every accepted pair.
// FILT: the gate runs through the single-precision filter (gjk.cuh: kdop_point_gate; same decisions); false = every axis in
// the reference arithmetic (TRAJOPT_B200_NP_FILTER=0, kept for A/B measurements)
template <bool FILT>
__device__ __forceinline__ bool np_gate(const NarrowArgs& a, uint32_t row, const double* s_kdop, const float* s_kdop_f, const double* pt,
                                        unsigned* groups, unsigned* exact, int axis_begin, int axis_end) {
  const double *lo = a.klo + (size_t)TOB_KDOP_AXES * row, *hi = a.khi + (size_t)TOB_KDOP_AXES * row;
  if (FILT) {
    const double2 c01 = *reinterpret_cast<const double2*>(a.kc + (size_t)4 * row);
    const double centre[3] = {c01.x, c01.y, a.kc[(size_t)4 * row + 2]};
    return kdop_point_gate(a.kf + (size_t)TOB_KF_ROW * row, centre, s_kdop_f, lo, hi, s_kdop, pt, a.dist, groups, exact, axis_begin, axis_end);
  }
  return kdop_point_overlap(lo, hi, s_kdop, pt, a.dist, groups, axis_begin, axis_end);
}

// MINB = resident CTAs per SM the register allocation is made for (4: 128 registers, 5: 102, 6: 85, 8: 64)
// PMEM: the hull vertices are read from memory in every GJK round (gjk.cuh: gjk_witness_6pt_mem) instead of held in registers
template <int MINB, bool FILT, bool PMEM = false>
__global__ void __launch_bounds__(NP_THREADS, MINB) k_narrow(NarrowArgs a) {
  __shared__ double s_kdop[3 * TOB_KDOP_AXES];
  __shared__ float s_kdop_f[FILT ? 3 * TOB_KDOP_AXES : 1];
  __shared__ uint32_t s_surv[NP_CHUNK];
  __shared__ uint32_t s_w[NP_THREADS / 32];
  for (int i = threadIdx.x; i < 3 * TOB_KDOP_AXES; i += blockDim.x) {
    s_kdop[i] = a.kdop[i];
    if (FILT) s_kdop_f[i] = (float)a.kdop[i];
  }
  const uint32_t n = a.dc->n_cand;
  if (n > a.cap) return;
  const uint32_t per = np_per_of(n, a.np_grid), chunk_sz = per * NP_THREADS;
  const uint32_t n_chunks = (n + chunk_sz - 1) / chunk_sz;
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t n_live = a.live_key ? a.dc->n_live : 0u;
  const int gate1 = a.gate1;
  unsigned w_groups = 0, w_iters = 0, w_exact = 0, w_band = 0;   // counted work of this thread
  __syncthreads();
  // chunks are handed out on demand: a chunk costs between a few gate tests and 512 GJK runs, and with ~20 chunks per CTA (the
  // shard one of eight GPUs gets of the batch) a fixed deal left the grid waiting for its unluckiest CTAs.  Which CTA runs a
  // chunk does not matter: planes and flags are stored by candidate index, the count by chunk.
  __shared__ uint32_t s_chunk;
  for (;;) {
    if (tid == 0) s_chunk = atomicAdd(&a.dc->np_next, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    if (chunk >= n_chunks) break;
    uint32_t n_surv = 0;   // uniform
    // the candidates of this thread are gathered first (index -> row / point -> coordinates are two dependent global loads
    // each; one after the other behind the block barriers of the compaction they cost four round trips instead of one)
    uint32_t c_row[NP_PER];
    double c_pt[NP_PER][3];
#pragma unroll
    for (uint32_t q = 0; q < NP_PER; q++) {
      const uint32_t i = chunk * chunk_sz + q * NP_THREADS + tid;
      c_row[q] = 0; c_pt[q][0] = c_pt[q][1] = c_pt[q][2] = 0.0;
      if (q < per && i < n) {
        const uint32_t p = a.cand_pt[i];
        c_row[q] = a.cand_row[i];
        c_pt[q][0] = a.px[p]; c_pt[q][1] = a.py[p]; c_pt[q][2] = a.pz[p];
      }
    }
    // the first axes of the gate, every candidate
#pragma unroll
    for (uint32_t q = 0; q < NP_PER; q++) {
      if (q >= per) break;   // uniform
      const uint32_t loc = q * NP_THREADS + tid, i = chunk * chunk_sz + loc;
      bool pass = false;
      if (i < n) {
        const uint32_t row = c_row[q];
        pass = np_gate<FILT>(a, row, s_kdop, s_kdop_f, c_pt[q], &w_groups, &w_exact, 0, gate1);
        a.cflag[i] = 0;
      }
      n_surv = np_append(pass, loc, s_surv, n_surv, s_w);
    }
    __syncthreads();
    // GJK + plane on the dense list of survivors
    uint32_t ok = 0;
    for (uint32_t sidx = tid; sidx < n_surv; sidx += NP_THREADS) {
      const uint32_t ii = chunk * chunk_sz + s_surv[sidx];
      const uint32_t row = a.cand_row[ii], p = a.cand_pt[ii];
      if (n_live) {   // is_seperate[tr_id][ob_id] (Optimization3D_admm.h:128): a live pair keeps its plane
        const unsigned long long key = ((unsigned long long)row << 32) | p;
        const uint32_t at = lower_bound_u64(a.live_key, n_live, key);
        if (at < n_live && a.live_key[at] == key) continue;
      }
      const double pt[3] = {a.px[p], a.py[p], a.pz[p]};
      double c[3], d, cn;
      if (PMEM) {
        gjk_witness_6pt_mem(a.P + (size_t)18 * row, pt, c, &w_iters);
        cn = eig_norm3(c);
      } else {
        double P[6][3];
        load_pts6(a.P + (size_t)18 * row, P);
        cn = plane_point_witness(P, pt, c, &w_iters);
      }
      bool acc = !(cn > a.dist);
      if (acc && !(cn <= a.gate_skip)) {      // inside the band (or NaN): the gate's own arithmetic decides
        w_band++;
        acc = np_gate<FILT>(a, row, s_kdop, s_kdop_f, pt, &w_groups, &w_exact, gate1, TOB_KDOP_AXES);
      }
      if (acc) {
        plane_point_finish(pt, a.offset, cn, c, &d);
        ok++;
        *reinterpret_cast<double4*>(a.cpl + (size_t)4 * ii) = make_double4(c[0], c[1], c[2], d);
        a.cflag[ii] = 1;
      }
    }
    // accepted planes of the chunk (integer sum: order-independent)
    for (int o = 16; o; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
    __syncthreads();
    if (lane == 0) s_w[w] = ok;
    __syncthreads();
    if (tid == 0) {
      uint32_t cnt = 0;
      for (int k = 0; k < NP_THREADS / 32; k++) cnt += s_w[k];
      a.csum[chunk] = cnt;
    }
  }
  for (int o = 16; o; o >>= 1) {
    w_groups += __shfl_xor_sync(0xffffffffu, w_groups, o);
    w_iters += __shfl_xor_sync(0xffffffffu, w_iters, o);
    w_exact += __shfl_xor_sync(0xffffffffu, w_exact, o);
    w_band += __shfl_xor_sync(0xffffffffu, w_band, o);
  }
  if (lane == 0 && w_groups) {
    atomicAdd(&a.dc->np_kdop_groups, (unsigned long long)w_groups);
    atomicAdd(&a.dc->np_gjk_iters, (unsigned long long)w_iters);
    if (w_exact) atomicAdd(&a.dc->np_kdop_exact, (unsigned long long)w_exact);
    if (w_band) atomicAdd(&a.dc->np_band, (unsigned long long)w_band);
  }
}

// ---- inter-robot planes ----------------------------------------------------------------------------------------
struct SelfArgs {
  int U, n_tr, npairs;
  const double *P, *D, *box, *klo, *khi, *kdop;
  double dist, offset, margin;
  double* self_pl;     // n_tr x npairs x 4
  uint32_t* self_ok;   // n_tr x npairs
  uint32_t* selfcnt;   // rows: accepted inter-robot planes per row (zeroed before the launch)
  // persistent mode (is_optimal_plane): live flag + plane per (slot, pair); nullptr otherwise
  uint32_t* live;
  double* lpl;
  DevCounts* dc;
};

__device__ __forceinline__ void pair_from_index(int idx, int U, int* p0, int* p1) {
  // lexicographic (p0<p1) enumeration
  int a = 0, rem = idx;
  while (rem >= U - 1 - a) { rem -= U - 1 - a; a++; }
  *p0 = a; *p1 = a + 1 + rem;
}

__global__ void __launch_bounds__(64) k_self_planes(SelfArgs a) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_tr * a.npairs) return;
  int tr = t / a.npairs, pi = t - tr * a.npairs, p0, p1;
  pair_from_index(pi, a.U, &p0, &p1);
  size_t r0 = (size_t)p0 * a.n_tr + tr, r1 = (size_t)p1 * a.n_tr + tr;
  const double *b0 = a.box + 6 * r0, *b1 = a.box + 6 * r1;
  uint32_t ok = 0;
  // aabb::Tree::query(margin) leaf test with node = p0, _node = p1 (AABB.cc:691,698): _box.overlaps(box, true, m)
  bool hit = true;
  for (int k = 0; k < 3; k++)
    if (b0[3 + k] + a.dist < b1[k] || b0[k] > b1[3 + k] + a.dist) hit = false;
  hit = hit && kdop_sets_overlap(a.klo + TOB_KDOP_AXES * r0, a.khi + TOB_KDOP_AXES * r0, a.klo + TOB_KDOP_AXES * r1,
                                 a.khi + TOB_KDOP_AXES * r1, a.dist);
  if (a.live) {
    // Optimization3D_multi.h:271-339: a pair that ever separated keeps its plane; every live plane is refined by
    // self_optimal_cd each iteration (also when the boxes no longer overlap) and always emitted.  Nothing is touched when
    // the iteration is going to be repeated (candidate overflow).
    if (a.dc->overflow) { a.self_ok[t] = 0; return; }
    uint32_t lv = a.live[t];
    if (!lv && !hit) { a.self_ok[t] = 0; return; }
    double P0[6][3], P1[6][3], c[3], d;
    load_pts6(a.P + 18 * r0, P0);
    load_pts6(a.P + 18 * r1, P1);
    double* st = a.lpl + (size_t)4 * t;
    if (!lv) {
      if (plane_hulls(P0, P1, a.dist, c, &d)) lv = 1;
    } else {
      c[0] = st[0]; c[1] = st[1]; c[2] = st[2]; d = st[3];
    }
    if (lv) {
      if (self_optimal_cd(P0, P1, a.offset, a.margin, c, &d)) atomicAdd(&a.dc->opt_capped, 1u);
      st[0] = c[0]; st[1] = c[1]; st[2] = c[2]; st[3] = d;
      a.live[t] = 1;
      double* o = a.self_pl + (size_t)4 * t;
      o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = d;
      atomicAdd(a.selfcnt + r0, 1u);
      atomicAdd(a.selfcnt + r1, 1u);
    }
    a.self_ok[t] = lv;
    return;
  }
  if (hit) {
    double P0[6][3], P1[6][3], c[3], d;
    load_pts6(a.P + 18 * r0, P0);
    load_pts6(a.P + 18 * r1, P1);
    if (plane_hulls(P0, P1, a.dist, c, &d)) {
      refine_d(P0, P1, c, a.offset, a.margin, &d, 10000);
      ok = 1;
      double* o = a.self_pl + (size_t)4 * t;
      o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = d;
      atomicAdd(a.selfcnt + r0, 1u);
      atomicAdd(a.selfcnt + r1, 1u);
    }
  }
  a.self_ok[t] = ok;
}

// ---- packing -----------------------------------------------------------------------------------------------------
// pair index of (lower, higher) robot in the lexicographic enumeration
__device__ __forceinline__ int pair_index(int a, int b, int U) { return a * (U - 1) - a * (a - 1) / 2 + (b - a - 1); }

struct PackArgs {
  DevCounts* dc;
  uint32_t cap;
  int rows_all, n_tr, U, npairs, with_self, self_begin, self_end;
  double offset;
  const uint32_t *cand_row, *cflag, *row_off, *selfcnt;
  uint32_t *csum, *selfpre;
  const uint32_t* self_ok;
  const double *cpl, *self_pl;
  double* pl;
  uint32_t *pl_row, *pl_off;
  int live;             // persistent-plane mode: the accepted planes are NEW members of the live set, not the plane list
  uint32_t live_cap;
  uint32_t np_grid;
};

// one CTA: scan of the per-chunk obstacle-plane counts and of the per-row inter-robot plane counts
__global__ void __launch_bounds__(1024) k_np_top(PackArgs a) {
  const uint32_t n = a.dc->n_cand;
  if (n > a.cap) return;
  const uint32_t chunk_sz = np_per_of(n, a.np_grid) * NP_THREADS;
  const uint32_t n_chunks = (n + chunk_sz - 1) / chunk_sz;
  const uint32_t ob_total = cta1024_scan_inplace(a.csum, n_chunks);
  for (int row = threadIdx.x; row < a.rows_all; row += 1024)
    a.selfpre[row] = (a.with_self && row >= a.self_begin && row < a.self_end) ? a.selfcnt[row] : 0u;
  __syncthreads();
  const uint32_t self_total = cta1024_scan_inplace(a.selfpre, (uint32_t)a.rows_all);
  if (threadIdx.x == 0) a.dc->n_en_items = 0;     // the listed parts of heavy rows are rebuilt from the new CSR (barrier.cu)
  if (threadIdx.x == 0) a.dc->np_next = 0;        // k_narrow has handed out its chunks: ready for the next plane pass
  if (threadIdx.x == 0 && a.live) {
    a.csum[n_chunks] = ob_total;
    a.selfpre[a.rows_all] = 0;
    a.dc->n_new = ob_total;
    if ((unsigned long long)a.dc->n_live + ob_total > a.live_cap) a.dc->overflow |= TOB_OVF_LIVE;
  } else if (threadIdx.x == 0) {
    a.csum[n_chunks] = ob_total;
    a.selfpre[a.rows_all] = self_total;
    a.dc->n_planes = ob_total + self_total;
    a.dc->n_planes_ob = ob_total;
    a.dc->planes += ob_total + self_total;
  }
}

// accepted obstacle planes before candidate idx (idx = a row boundary): scanned chunk base + the flags of the chunk in
// front of idx, summed by the warp (uniform result)
__device__ __forceinline__ uint32_t ob_prefix_warp(const PackArgs& a, uint32_t idx, uint32_t n, uint32_t n_chunks, uint32_t chunk_sz) {
  if (idx >= n) return a.csum[n_chunks];
  const uint32_t chunk = idx / chunk_sz, lane = threadIdx.x & 31;
  uint32_t local = 0;
  for (uint32_t j = chunk * chunk_sz + lane; j < idx; j += 32) local += a.cflag[j];
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  return a.csum[chunk] + local;
}

// scatter: obstacle planes keep the candidate order (row, Morton position); the inter-robot planes of a row go behind
// its obstacle planes: (c, d-offset/2) for the lower robot id, (-c, -d-offset/2) for the higher one
// (Optimization3D_multi.h:300-304).  Also writes the CSR offsets pl_off[rows_all+1].
__global__ void __launch_bounds__(NP_THREADS) k_pack(PackArgs a) {
  __shared__ uint32_t s_w[NP_THREADS / 32];
  const uint32_t n = a.dc->n_cand;
  if (n > a.cap) return;
  const uint32_t per = np_per_of(n, a.np_grid), chunk_sz = per * NP_THREADS;
  const uint32_t n_chunks = (n + chunk_sz - 1) / chunk_sz;
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    uint32_t run = a.csum[chunk];     // uniform: planes before this pass of the chunk
#pragma unroll
    for (uint32_t q = 0; q < NP_PER; q++) {
      if (q >= per) break;   // uniform
      const uint32_t i = chunk * chunk_sz + q * NP_THREADS + tid;
      const bool f = i < n && a.cflag[i];
      const uint32_t bm = __ballot_sync(0xffffffffu, f);
      __syncthreads();
      if (lane == 0) s_w[w] = __popc(bm);
      __syncthreads();
      uint32_t rank = __popc(bm & ((1u << lane) - 1u)), tot = 0;
#pragma unroll
      for (int k = 0; k < NP_THREADS / 32; k++) {
        if (k < (int)w) rank += s_w[k];
        tot += s_w[k];
      }
      if (f) {
        const uint32_t row = a.cand_row[i];
        const uint32_t dst = run + rank + a.selfpre[row];
        const double4 v = *reinterpret_cast<const double4*>(a.cpl + (size_t)4 * i);
        *reinterpret_cast<double4*>(a.pl + (size_t)4 * dst) = v;
        a.pl_row[dst] = row;
      }
      run += tot;
    }
  }
  // CSR offsets and inter-robot planes: one warp per row
  const int n_warps = gridDim.x * (NP_THREADS / 32);
  for (int row = blockIdx.x * (NP_THREADS / 32) + w; row <= a.rows_all; row += n_warps) {
    const uint32_t pre0 = ob_prefix_warp(a, a.row_off[row], n, n_chunks, chunk_sz);
    const uint32_t off = pre0 + a.selfpre[row];
    if (lane == 0) a.pl_off[row] = off;
    if (row < a.rows_all && a.with_self && row >= a.self_begin && row < a.self_end) {
      const uint32_t pre1 = ob_prefix_warp(a, a.row_off[row + 1], n, n_chunks, chunk_sz);
      uint32_t dst = off + (pre1 - pre0);
      const int u = row / a.n_tr, tr = row - u * a.n_tr;
      for (int v0 = 0; v0 < a.U; v0 += 32) {
        const int v = v0 + lane;
        size_t t = 0;
        bool okv = false;
        int lo = 0;
        if (v < a.U && v != u) {
          lo = v < u ? v : u;
          const int hi = v < u ? u : v;
          t = (size_t)tr * a.npairs + pair_index(lo, hi, a.U);
          okv = a.self_ok[t] != 0;
        }
        const uint32_t bm = __ballot_sync(0xffffffffu, okv);
        if (okv) {
          const uint32_t at = dst + __popc(bm & ((1u << lane) - 1u));
          const double* sp = a.self_pl + 4 * t;
          double* o = a.pl + (size_t)4 * at;
          if (u == lo) { o[0] = sp[0]; o[1] = sp[1]; o[2] = sp[2]; o[3] = sp[3] - 0.5 * a.offset; }
          else { o[0] = -sp[0]; o[1] = -sp[1]; o[2] = -sp[2]; o[3] = -sp[3] - 0.5 * a.offset; }
          a.pl_row[at] = row;
        }
        dst += __popc(bm);
      }
    }
  }
}

// ---- persistent planes -------------------------------------------------------------------------------------------
struct LiveArgs {
  DevCounts* dc;
  uint32_t cap;          // candidate capacity
  int rows_all;
  const uint32_t *cand_pt, *cand_row, *cflag, *csum;
  const double* cpl;
  unsigned long long *live_key, *tmp_key, *new_key;
  double *live_pl, *tmp_pl, *new_pl;
  const double *px, *py, *pz, *P;
  double offset, margin;
  double* pl;
  uint32_t *pl_row, *pl_off;
  uint32_t np_grid;
};

// planes accepted for pairs that were not live: compacted in candidate order = (row, Morton position) = key order
__global__ void __launch_bounds__(NP_THREADS) k_live_compact(LiveArgs a) {
  __shared__ uint32_t s_w[NP_THREADS / 32];
  const uint32_t n = a.dc->n_cand;
  if (a.dc->overflow) return;
  const uint32_t per = np_per_of(n, a.np_grid), chunk_sz = per * NP_THREADS;
  const uint32_t n_chunks = (n + chunk_sz - 1) / chunk_sz;
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    uint32_t run = a.csum[chunk];
#pragma unroll
    for (uint32_t q = 0; q < NP_PER; q++) {
      if (q >= per) break;   // uniform
      const uint32_t i = chunk * chunk_sz + q * NP_THREADS + tid;
      const bool f = i < n && a.cflag[i];
      const uint32_t bm = __ballot_sync(0xffffffffu, f);
      __syncthreads();
      if (lane == 0) s_w[w] = __popc(bm);
      __syncthreads();
      uint32_t rank = __popc(bm & ((1u << lane) - 1u)), tot = 0;
#pragma unroll
      for (int k = 0; k < NP_THREADS / 32; k++) {
        if (k < (int)w) rank += s_w[k];
        tot += s_w[k];
      }
      if (f) {
        const uint32_t dst = run + rank;
        a.new_key[dst] = ((unsigned long long)a.cand_row[i] << 32) | a.cand_pt[i];
        *reinterpret_cast<double4*>(a.new_pl + (size_t)4 * dst) = *reinterpret_cast<const double4*>(a.cpl + (size_t)4 * i);
      }
      run += tot;
    }
  }
}

// merge of the two sorted, disjoint key sets (live, new) into tmp: every element finds its place by one binary search
__global__ void __launch_bounds__(256) k_live_merge(LiveArgs a) {
  if (a.dc->overflow) return;
  const uint32_t n_old = a.dc->n_live, n_new = a.dc->n_new;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_old + n_new; idx += stride) {
    unsigned long long key;
    uint32_t pos;
    const double* src;
    if (idx < n_old) {
      key = a.live_key[idx];
      pos = idx + lower_bound_u64(a.new_key, n_new, key);
      src = a.live_pl + (size_t)4 * idx;
    } else {
      const uint32_t j = idx - n_old;
      key = a.new_key[j];
      pos = j + lower_bound_u64(a.live_key, n_old, key);
      src = a.new_pl + (size_t)4 * j;
    }
    a.tmp_key[pos] = key;
    *reinterpret_cast<double4*>(a.tmp_pl + (size_t)4 * pos) = *reinterpret_cast<const double4*>(src);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.dc->n_live_next = n_old + n_new;
}

// Optimal_plane::optimal_cd on every live plane (Optimization3D_admm.h:164-193), one plane per thread; the refined planes
// are both the persistent state (live_*) and the plane list of this iteration (pl / pl_row / pl_off, CSR over rows)
__global__ void __launch_bounds__(128) k_live_refine(LiveArgs a) {
  if (a.dc->overflow) return;
  const uint32_t n = a.dc->n_live_next;
  const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t i = t0; i < n; i += stride) {
    const unsigned long long key = a.tmp_key[i];
    const uint32_t row = (uint32_t)(key >> 32), p = (uint32_t)(key & 0xffffffffu);
    double4 v = *reinterpret_cast<const double4*>(a.tmp_pl + (size_t)4 * i);
    const double q[3] = {a.px[p], a.py[p], a.pz[p]};
    double P[6][3], c[3] = {v.x, v.y, v.z}, d = v.w;
    load_pts6(a.P + (size_t)18 * row, P);
    if (optimal_cd(P, q, a.offset, a.margin, c, &d)) atomicAdd(&a.dc->opt_capped, 1u);
    v = make_double4(c[0], c[1], c[2], d);
    a.live_key[i] = key;
    *reinterpret_cast<double4*>(a.live_pl + (size_t)4 * i) = v;
    *reinterpret_cast<double4*>(a.pl + (size_t)4 * i) = v;
    a.pl_row[i] = row;
  }
  for (uint32_t row = t0; row <= (uint32_t)a.rows_all; row += stride)
    a.pl_off[row] = lower_bound_u64(a.tmp_key, n, (unsigned long long)row << 32);
  if (t0 == 0) {
    a.dc->n_live = n;
    a.dc->n_planes = n;
    a.dc->n_planes_ob = n;
    a.dc->planes += n;
  }
}

// (re)allocates the persistent set for `need` planes, keeping the live ones; also sizes the plane list for them
int ensure_live_buffers(tob_ctx* c, uint64_t need) {
  if (c->live_cap >= need && c->live_key.p) return 0;
  uint64_t cap = c->live_cap;
  if (!cap) {
    cap = 1u << 20;
    if (const char* e = getenv("TRAJOPT_B200_LIVE_CAP")) { long v = atol(e); if (v >= 16) cap = (uint64_t)v; }   // tests: force growth
  }
  while (cap < need) cap *= 2;
  if (cap > 0xfff00000ull) return fail_msg(c, "live plane count exceeds the 32-bit index range");
  uint32_t n_live = 0;
  if (c->live_key.p && c->dc.p) {
    TOB_CUDA(c, cudaStreamSynchronize(c->stream));
    TOB_CUDA(c, cudaMemcpy(&n_live, &c->dc.p->n_live, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  DBuf<unsigned long long> nk;
  DBuf<double> np;
  TOB_CUDA(c, nk.ensure(cap + 1));
  TOB_CUDA(c, np.ensure(4 * cap + 4));
  if (n_live) {
    TOB_CUDA(c, cudaMemcpy(nk.p, c->live_key.p, (size_t)n_live * sizeof(unsigned long long), cudaMemcpyDeviceToDevice));
    TOB_CUDA(c, cudaMemcpy(np.p, c->live_pl.p, (size_t)4 * n_live * sizeof(double), cudaMemcpyDeviceToDevice));
  }
  c->live_key = std::move(nk); c->live_pl = std::move(np);
  alloc_generation()++;
  TOB_CUDA(c, c->live_key_t.ensure(cap + 1)); TOB_CUDA(c, c->live_pl_t.ensure(4 * cap + 4));
  TOB_CUDA(c, c->new_key.ensure(cap + 1)); TOB_CUDA(c, c->new_pl.ensure(4 * cap + 4));
  c->live_cap = cap;
  return ensure_query_buffers(c);
}

int reset_live_planes(tob_ctx* c) {
  if (c->dc.p) TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->n_live, 0, 4 * sizeof(uint32_t), c->stream));
  if (c->self_live.p) TOB_CUDA(c, cudaMemsetAsync(c->self_live.p, 0, c->self_live.cap * sizeof(uint32_t), c->stream));
  return 0;
}

// tail of the plane pass in persistent mode: new planes -> compact -> merge with the live set -> refine all -> CSR
static int live_rows(tob_ctx* c) {
  cudaStream_t st = c->stream;
  LiveArgs a;
  a.dc = c->dc.p; a.cap = (uint32_t)c->cand_cap; a.rows_all = c->rows_all();
  a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p; a.cflag = c->cflag.p; a.csum = c->csum.p; a.cpl = c->cpl.p;
  a.live_key = c->live_key.p; a.tmp_key = c->live_key_t.p; a.new_key = c->new_key.p;
  a.live_pl = c->live_pl.p; a.tmp_pl = c->live_pl_t.p; a.new_pl = c->new_pl.p;
  a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p; a.P = c->geo.P.p;
  a.offset = c->prm.offset; a.margin = c->prm.margin;
  a.pl = c->pl.p; a.pl_row = c->pl_row.p; a.pl_off = c->pl_off.p;
  a.np_grid = np_grid_of(c);
  {
    Prof prof(c, K_PACK);
    k_live_compact<<<c->sm_count * 4, NP_THREADS, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  {
    Prof prof(c, K_PACK);
    k_live_merge<<<c->sm_count * 4, 256, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  {
    Prof prof(c, K_LIVE_REFINE);
    k_live_refine<<<c->sm_count * 8, 128, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

int self_planes(tob_ctx* c) {
  const int U = c->n_robots();
  int npairs = U * (U - 1) / 2;
  size_t n = (size_t)c->n_tr * npairs;
  TOB_CUDA(c, c->self_pl.ensure(4 * n + 4));
  TOB_CUDA(c, c->self_ok.ensure(n + 1));
  SelfArgs a;
  a.U = U; a.n_tr = c->n_tr; a.npairs = npairs;
  a.P = c->geo.P.p; a.D = nullptr; a.box = c->geo.box.p; a.klo = c->geo.klo.p; a.khi = c->geo.khi.p; a.kdop = c->d_kdop.p;
  a.dist = c->prm.offset + 2 * c->prm.margin; a.offset = c->prm.offset; a.margin = c->prm.margin;
  a.self_pl = c->self_pl.p; a.self_ok = c->self_ok.p; a.selfcnt = c->selfcnt.p;
  a.live = nullptr; a.lpl = nullptr; a.dc = c->dc.p;
  if (c->prm.optimal_plane) {
    if (!c->self_live.p || c->self_live.cap < n + 1) {     // first use (or more pairs): all pairs start non-live
      TOB_CUDA(c, c->self_live.ensure(n + 1));
      TOB_CUDA(c, c->self_lpl.ensure(4 * n + 4));
      TOB_CUDA(c, cudaMemsetAsync(c->self_live.p, 0, c->self_live.cap * sizeof(uint32_t), c->stream));
    }
    a.live = c->self_live.p; a.lpl = c->self_lpl.p;
  }
  TOB_CUDA(c, cudaMemsetAsync(c->selfcnt.p, 0, (size_t)c->rows_all() * sizeof(uint32_t), c->stream));
  if (n) {
    Prof prof(c, K_SELF_PLANES);
    k_self_planes<<<div_up(n, 64), 64, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  c->ctr.self_pairs += n;
  return 0;
}

// shared tail: chunk / row scans -> scatter into the CSR.  Candidate flags (cflag, csum) and row_off must be current.
static int pack_rows(tob_ctx* c, int rb, int re, bool ws, bool live = false) {
  cudaStream_t st = c->stream;
  const int U = c->n_robots();
  TOB_CUDA(c, c->self_ok.ensure(1));
  TOB_CUDA(c, c->self_pl.ensure(4));
  PackArgs a;
  a.dc = c->dc.p; a.cap = (uint32_t)c->cand_cap;
  a.rows_all = c->rows_all(); a.n_tr = c->n_tr; a.U = U; a.npairs = U * (U - 1) / 2; a.with_self = ws ? 1 : 0;
  a.self_begin = rb * c->n_tr; a.self_end = re * c->n_tr; a.offset = c->prm.offset;
  a.cand_row = c->cand_row.p; a.cflag = c->cflag.p; a.row_off = c->row_off.p; a.csum = c->csum.p; a.selfpre = c->selfpre.p;
  a.selfcnt = c->selfcnt.p;
  a.self_ok = c->self_ok.p; a.cpl = c->cpl.p; a.self_pl = c->self_pl.p;
  a.pl = c->pl.p; a.pl_row = c->pl_row.p; a.pl_off = c->pl_off.p;
  a.live = live ? 1 : 0; a.live_cap = (uint32_t)c->live_cap; a.np_grid = np_grid_of(c);
  {
    Prof prof(c, K_SCAN);
    k_np_top<<<1, 1024, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  if (live) { TOB_TRY(live_rows(c)); return energy_items(c, false); }
  {
    Prof prof(c, K_PACK);
    // chunks as k_narrow cut them (np_grid), but spread over every resident CTA slot: the scatter is a chain of dependent loads
    // (flag -> row -> plane) per thread, bound by the number of warps in flight (ncu: 33 cycles of long-scoreboard stall per issue
    // at 4 CTAs per SM, 0.33 of the HBM peak; 16 per SM: 0.182 -> 0.089 ms on a 128-problem shard of the batch)
    const char* e = getenv("TRAJOPT_B200_PACK_GRID");
    int per_sm = e ? atoi(e) : 16;
    if (per_sm < 1 || per_sm > 32) per_sm = 16;
    k_pack<<<c->sm_count * per_sm, NP_THREADS, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  return energy_items(c, false);   // k_np_top zeroed the counter
}

// candidates (c->cand_*, c->row_off, dc->n_cand) + row geometry must be current.  Leaves the packed plane CSR.
// Inter-robot planes need geo.P/box/klo/khi of ALL robots.  Asynchronous (no host read-back).
int narrowphase_planes(tob_ctx* c, int rb, int re, int with_self) {
  cudaStream_t st = c->stream;
  TOB_TRY(ensure_query_buffers(c));
  NarrowArgs a;
  a.dc = c->dc.p; a.cap = (uint32_t)c->cand_cap; a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p;
  a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p;
  a.P = c->geo.P.p; a.klo = c->geo.klo.p; a.khi = c->geo.khi.p; a.kdop = c->d_kdop.p;
  a.kf = c->geo.kf.p; a.kc = c->geo.kc.p;
  a.dist = c->prm.offset + c->prm.margin; a.offset = c->prm.offset;     // the filter's thresholds are made for this gap (segments.cu)
  a.cpl = c->cpl.p; a.cflag = c->cflag.p; a.csum = c->csum.p;
  const bool live = c->live_planes();
  if (live) {
    if (rb != 0 || re != c->n_robots()) return fail_msg(c, "persistent planes: the plane pass must cover every robot slot of the context");
    TOB_TRY(ensure_live_buffers(c, c->live_cap ? c->live_cap : 1));
  }
  a.live_key = live ? c->live_key.p : nullptr;
  a.np_grid = np_grid_of(c);
  {
    Prof prof(c, K_NARROW);
    const char *eo = getenv("TRAJOPT_B200_NP_OCC"), *ef = getenv("TRAJOPT_B200_NP_FILTER");   // A/B measurements, tests
    const char *eb = getenv("TRAJOPT_B200_NP_BAND"), *eg = getenv("TRAJOPT_B200_NP_GATE1");
    const int occ = eo ? atoi(eo) : 4, filt = ef ? atoi(ef) : 1;
    a.gate_skip = (eb && atoi(eb) == 0) ? -1.0 : a.dist * (1.0 - 1e-6);
    a.gate1 = eg ? atoi(eg) : NP_GATE1;
    if (a.gate1 < 7 || a.gate1 > TOB_KDOP_AXES || a.gate1 % 7) a.gate1 = NP_GATE1;
    a.np_grid = np_grid_of(c);
    // chunk ticket of k_narrow: k_np_top leaves it at zero, but a plane pass that was abandoned between the two (an error
    // return) must not starve the next one
    TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->np_next, 0, sizeof(uint32_t), st));
    const char* em = getenv("TRAJOPT_B200_NP_PMEM");
    // default: hull vertices from L1, registers for 5 CTAs per SM (measured on a 128-problem shard: 1.025 ms with the vertices
    // in registers at 4 CTAs per SM, 1.001 / 0.959 / 0.990 ms with PMEM at 4 / 5 / 6); 0 = the register variants below
    const int pmem = em ? atoi(em) : 5;
    if (pmem == 4) k_narrow<4, false, true><<<c->sm_count * 4, NP_THREADS, 0, st>>>(a);
    else if (pmem == 5) k_narrow<5, false, true><<<c->sm_count * 5, NP_THREADS, 0, st>>>(a);
    else if (pmem == 6) k_narrow<6, false, true><<<c->sm_count * 6, NP_THREADS, 0, st>>>(a);
    else if (!filt) k_narrow<4, false><<<c->sm_count * 4, NP_THREADS, 0, st>>>(a);
    else if (occ == 8) k_narrow<8, true><<<c->sm_count * 8, NP_THREADS, 0, st>>>(a);
    else if (occ == 6) k_narrow<6, true><<<c->sm_count * 6, NP_THREADS, 0, st>>>(a);
    else if (occ == 5) k_narrow<5, true><<<c->sm_count * 5, NP_THREADS, 0, st>>>(a);
    else k_narrow<4, true><<<c->sm_count * 4, NP_THREADS, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  if (with_self < 0) return 0;      // obstacle part only: the caller finishes with narrowphase_finish()
  bool ws = with_self && c->n_robots() > 1;
  if (ws) TOB_TRY(self_planes(c));
  return pack_rows(c, rb, re, ws, live);
}

// second half of the plane pass when the first was launched with with_self < 0: inter-robot planes (need geo of ALL
// robots), then the packing of both kinds into the CSR
int narrowphase_finish(tob_ctx* c, int rb, int re, int with_self) {
  bool ws = with_self && c->n_robots() > 1;
  if (ws) TOB_TRY(self_planes(c));
  return pack_rows(c, rb, re, ws, c->live_planes());
}

// inter-robot planes only (Optimization3D_multi::separate_self on empty lists): geo of ALL robots must be current
int pack_self_only(tob_ctx* c) {
  const int rows = c->rows_all();
  TOB_TRY(ensure_query_buffers(c));
  TOB_CUDA(c, cudaMemsetAsync(c->row_off.p, 0, (rows + 2) * sizeof(uint32_t), c->stream));
  TOB_CUDA(c, cudaMemsetAsync(&c->dc.p->n_cand, 0, sizeof(uint32_t), c->stream));
  TOB_TRY(self_planes(c));
  return pack_rows(c, 0, c->n_robots(), true);
}

// caller-provided plane lists (the reference's c_lists/d_lists) for robots [rb,re): offsets over (re-rb)*n_tr rows
int pack_planes_from_host(tob_ctx* c, int rb, int re, const uint32_t* offsets, const double* cc, const double* dd) {
  cudaStream_t st = c->stream;
  const int rows = c->rows_all(), nloc = (re - rb) * c->n_tr;
  uint64_t np = offsets[nloc];
  if (np > c->cand_cap) TOB_TRY(grow_cand_capacity(c, np));
  TOB_TRY(ensure_query_buffers(c));
  std::vector<uint32_t> off(rows + 1), prow(np + 1);
  std::vector<double> pl(4 * np + 4);
  for (int g = 0; g <= rows; g++) {
    int lr = g - rb * c->n_tr;
    off[g] = lr < 0 ? 0u : (lr > nloc ? (uint32_t)np : offsets[lr]);
  }
  for (int lr = 0; lr < nloc; lr++)
    for (uint32_t k = offsets[lr]; k < offsets[lr + 1]; k++) {
      pl[4 * k] = cc[3 * k]; pl[4 * k + 1] = cc[3 * k + 1]; pl[4 * k + 2] = cc[3 * k + 2]; pl[4 * k + 3] = dd[k];
      prow[k] = rb * c->n_tr + lr;
    }
  uint32_t np32[2] = {(uint32_t)np, (uint32_t)np};
  TOB_CUDA(c, cudaMemcpyAsync(c->pl_off.p, off.data(), (rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(c->pl.p, pl.data(), (4 * np + 4) * sizeof(double), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(c->pl_row.p, prow.data(), (np + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(&c->dc.p->n_planes, np32, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TOB_TRY(energy_items(c, true));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  c->n_planes = np;
  return 0;
}

// ---- CCD ladder vs the cloud -----------------------------------------------------------------------------------
// Fused with the swept-box traversal (bp.cuh): the line search needs only max_k over all (sub-segment, point) pairs of
// the 0.8^k ladder exponent, not a candidate list, so every lane whose point passes the leaf predicate runs the CCD
// test in place (swept 49-DOP gate, then GJK on hull(P U P+sD) down the ladder) and atomicMax-es the exponent of its
// robot.  No count / scan / fill, no candidate buffer, no host read-back.
struct CcdArgs {
  BpArgs bp;
  const double *P, *D, *kdop, *steps;
  double offset;
  int n_tr;
  int* kmax;   // per robot
};

__device__ __forceinline__ void moved_points(const double (*P)[3], const double (*D)[3], double s, double (*A)[3]) {
  for (int j = 0; j < 6; j++)
    for (int k = 0; k < 3; k++) {
      A[j][k] = P[j][k] + 0.0 * D[j][k];
      A[j + 6][k] = P[j][k] + s * D[j][k];
    }
}

// MINB: resident CTAs per SM the registers are allotted for.  The traversal needs few registers, the ladder (swept 49-DOP +
// GJK on 12 + 1 points) many; with thousands of rows nearly every CTA only traverses (377 k of 73 M point tests reach the ladder
// on the 1024-problem batch), so the many-row launch takes the variant with more resident warps and lets the rare ladder spill.
template <int MINB>
__global__ void __launch_bounds__(BP_THREADS, MINB) k_bp_ccd(CcdArgs a) {
  __shared__ BpShared s;
  __shared__ double s_kdop[3 * TOB_KDOP_AXES];
  for (int i = threadIdx.x; i < 3 * TOB_KDOP_AXES; i += blockDim.x) s_kdop[i] = a.kdop[i];
  uint32_t rank;
  const uint32_t n_items = bp_prepare(a.bp, s, &rank);   // contains barriers: s_kdop is visible afterwards
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t cnt = 0;
  unsigned w_iters = 0, w_pass = 0;
  uint32_t h = 0;                                          // bp_item cursor
  for (uint32_t j = w; j < n_items; j += BP_WARPS) {
    uint32_t row, leaf;
    bp_item(a.bp, s, j, &h, &row, &leaf);
    double q[1][3];
    const bool ok = bp_point_test(a.bp, s, h, leaf * 32 + lane, &q[0][0], &q[0][1], &q[0][2]);
    cnt += __popc(__ballot_sync(0xffffffffu, ok));
    if (!ok) continue;
    const int robot = row / a.n_tr;
    double P[6][3], D[6][3], A[12][3];
    load_pts6(a.P + (size_t)18 * row, P);
    load_pts6(a.D + (size_t)18 * row, D);
    // Step::position_step carries `step` across pairs; the swept hull shrinks monotonically with the step, so starting
    // from the current exponent of the robot (a racy but always valid lower bound) gives the same result in any order
    int k = *((volatile int*)(a.kmax + robot));
    double st = a.steps[k];
    moved_points(P, D, st, A);
    // CCD::KDOPCCD(P, D, q, offset, 0, step)
    if (!kdop_overlap<12, 1>(A, q, s_kdop, a.offset)) continue;
    w_pass++;
    const double d2 = a.offset * a.offset;
    while (k < TOB_MAX_LADDER) {
      double v[3];
      gjk_witness<12, 1>(A, q, v, &w_iters);
      const double dist2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (!(dist2 <= d2)) break;
      k++;
      st = a.steps[k];
      moved_points(P, D, st, A);
    }
    atomicMax(a.kmax + robot, k);
  }
  if (lane == 0 && cnt) atomicAdd(&a.bp.dc->ccd_candidates, (unsigned long long)cnt);
  w_iters = __reduce_add_sync(0xffffffffu, w_iters);
  w_pass = __reduce_add_sync(0xffffffffu, w_pass);
  if (lane == 0 && w_pass) {
    atomicAdd(&a.bp.dc->ccd_gjk_iters, (unsigned long long)w_iters);
    atomicAdd(&a.bp.dc->ccd_kdop_pass, (unsigned long long)w_pass);
  }
}

// geo.P / geo.D / geo.box (swept) of robots [rb,re) must be current (compute_rows mode 3, which also zeroes kmax).
int ccd_position_steps(tob_ctx* c, int rb, int re) {
  if (c->n_pts == 0) return fail_msg(c, "CCD: no point cloud uploaded");
  TOB_TRY(ensure_query_buffers(c));
  CcdArgs a;
  bp_args(c, rb * c->n_tr, (re - rb) * c->n_tr, c->prm.offset, a.bp);
  a.P = c->geo.P.p; a.D = c->geo.D.p; a.kdop = c->d_kdop.p; a.steps = c->d_steps.p; a.offset = c->prm.offset; a.n_tr = c->n_tr;
  a.kmax = c->kmax.p;
  if (a.bp.n_tasks) {
    Prof prof(c, K_CCD);
    const char* e = getenv("TRAJOPT_B200_CCD_OCC");
    const int occ = e ? atoi(e) : 0;
    const int grid = div_up((size_t)a.bp.n_tasks, a.bp.tpc);
    const int use = occ ? occ : (a.bp.rows >= 8192 ? 8 : 4);
    if (use >= 12) k_bp_ccd<12><<<grid, BP_THREADS, 0, c->stream>>>(a);
    else if (use >= 8) k_bp_ccd<8><<<grid, BP_THREADS, 0, c->stream>>>(a);
    else k_bp_ccd<4><<<grid, BP_THREADS, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

// ---- inter-robot CCD ---------------------------------------------------------------------------------------------
struct SelfCcdArgs {
  int U, n_tr, npairs, coupled;
  const double *P, *D, *kdop, *steps;
  double offset;
  uint32_t* hit;      // list of task ids (slot*npairs + pair) that collide at full steps (1,1); hit[cap] = count
  uint32_t cap;
  int* kmax;          // [0]: shared exponent (coupled)
};

__device__ __forceinline__ void swept12(const double (*P)[3], const double (*D)[3], double t0, double t1, double (*A)[3]) {
  for (int j = 0; j < 6; j++)
    for (int k = 0; k < 3; k++) {
      A[j][k] = P[j][k] + t0 * D[j][k];
      A[j + 6][k] = P[j][k] + t1 * D[j][k];
    }
}

__device__ __forceinline__ bool self_swept_box_hit(const double (*P0)[3], const double (*D0)[3], const double (*P1)[3],
                                                   const double (*D1)[3], double d) {
  // BVH::SelfCCDCollision boxes (BVH.cpp:300-321): min/max over P and (P+D), full step
  for (int k = 0; k < 3; k++) {
    double lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
    for (int j = 0; j < 6; j++) {
      double v = P0[j][k], w = P0[j][k] + D0[j][k];
      if (v < lo0) lo0 = v; if (v > hi0) hi0 = v;
      if (w < lo0) lo0 = w; if (w > hi0) hi0 = w;
      v = P1[j][k]; w = P1[j][k] + D1[j][k];
      if (v < lo1) lo1 = v; if (v > hi1) hi1 = v;
      if (w < lo1) lo1 = w; if (w > hi1) hi1 = w;
    }
    if (hi0 + d < lo1 || lo0 > hi1 + d) return false;
  }
  return true;
}

// phase 1 (parallel): which (slot, pair) collide when both robots take their full step
__global__ void __launch_bounds__(64) k_self_ccd_filter(SelfCcdArgs a) {
  __shared__ double s_kdop[3 * TOB_KDOP_AXES];
  for (int i = threadIdx.x; i < 3 * TOB_KDOP_AXES; i += blockDim.x) s_kdop[i] = a.kdop[i];
  __syncthreads();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_tr * a.npairs) return;
  int tr = t / a.npairs, pi = t - tr * a.npairs, p0, p1;
  pair_from_index(pi, a.U, &p0, &p1);
  size_t r0 = (size_t)p0 * a.n_tr + tr, r1 = (size_t)p1 * a.n_tr + tr;
  double P0[6][3], D0[6][3], P1[6][3], D1[6][3];
  load_pts6(a.P + 18 * r0, P0); load_pts6(a.D + 18 * r0, D0);
  load_pts6(a.P + 18 * r1, P1); load_pts6(a.D + 18 * r1, D1);
  uint32_t hit = 0;
  if (self_swept_box_hit(P0, D0, P1, D1, a.offset)) {
    double A[12][3], B[12][3];
    swept12(P0, D0, 0.0, 1.0, A);
    swept12(P1, D1, 0.0, 1.0, B);
    if (kdop_overlap<12, 12>(A, B, s_kdop, a.offset)) {
      double v[3];
      gjk_witness<12, 12>(A, B, v);
      double dist2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (dist2 <= a.offset * a.offset) hit = 1;
    }
  }
  if (hit) {
    uint32_t k = atomicAdd(a.hit + a.cap, 1u);
    if (k < a.cap) a.hit[k] = (uint32_t)t;
  }
}

// phase 2 (one warp; the resolution itself is sequential like the reference): resolve the colliding pairs in (slot, pair)
// order.  The list is short (pairs that really collide when both robots take their full Newton step) and can never
// overflow: its capacity is the number of (slot, pair) tasks.  It is sorted first so the result does not depend on the
// order the filter threads appended it: every lane ranks its elements by counting the smaller ones (ids are unique).
__global__ void __launch_bounds__(32) k_self_ccd_resolve(SelfCcdArgs a, double* steps_out, uint32_t* sorted) {
  const uint32_t lane = threadIdx.x;
  const uint32_t nh = min(a.hit[a.cap], a.cap);
  for (uint32_t i = lane; i < nh; i += 32) {
    const uint32_t v = a.hit[i];
    uint32_t r = 0;
    for (uint32_t j = 0; j < nh; j++) r += a.hit[j] < v;
    sorted[r] = v;
  }
  __syncwarp();
  if (lane != 0) return;
  for (int u = 0; u < a.U; u++) steps_out[u] = 1.0;
  int kshared = 0;
  for (uint32_t i = 0; i < nh; i++) {
    const int tr = sorted[i] / a.npairs, pi = sorted[i] - tr * a.npairs;
    int p0, p1;
    pair_from_index(pi, a.U, &p0, &p1);
    size_t r0 = (size_t)p0 * a.n_tr + tr, r1 = (size_t)p1 * a.n_tr + tr;
    double P0[6][3], D0[6][3], P1[6][3], D1[6][3], A[12][3], B[12][3];
    load_pts6(a.P + 18 * r0, P0); load_pts6(a.D + 18 * r0, D0);
    load_pts6(a.P + 18 * r1, P1); load_pts6(a.D + 18 * r1, D1);
    double s0 = a.coupled ? a.steps[kshared] : steps_out[p0];
    double s1 = a.coupled ? a.steps[kshared] : steps_out[p1];
    swept12(P0, D0, 0.0, s0, A);
    swept12(P1, D1, 0.0, s1, B);
    if (!kdop_overlap<12, 12>(A, B, a.kdop, a.offset)) continue;
    int guard = 0;
    // the ladder table has TOB_LADDER + 2 entries: a shared exponent that reaches TOB_LADDER stays there (0.8^400 ~ 1e-39:
    // two robots that are already closer than `offset` never separate; the reference would loop forever)
    while (guard++ < TOB_MAX_LADDER && kshared < TOB_LADDER) {
      double v[3];
      gjk_witness<12, 12>(A, B, v);
      double dist2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (!(dist2 <= a.offset * a.offset)) break;
      if (a.coupled) { kshared++; s0 = s1 = a.steps[kshared]; }
      else { s0 *= 0.8; s1 *= 0.8; }
      swept12(P0, D0, 0.0, s0, A);
      swept12(P1, D1, 0.0, s1, B);
    }
    if (!a.coupled) { steps_out[p0] = s0; steps_out[p1] = s1; }
  }
  if (a.coupled) steps_out[0] = a.steps[kshared];
}

// geo.P / geo.D of ALL robots must be current (compute_rows with mode 2).  steps_dev: n_robots (or [0] coupled).
int self_ccd_steps(tob_ctx* c, int coupled, double* steps_dev) {
  const int U = c->n_robots();
  int npairs = U * (U - 1) / 2;
  size_t n = (size_t)c->n_tr * npairs;
  const uint32_t cap = (uint32_t)(n ? n : 1);   // every task can be listed: the list cannot overflow
  TOB_CUDA(c, c->self_hits.ensure(2 * (size_t)cap + 2));
  SelfCcdArgs a;
  a.U = U; a.n_tr = c->n_tr; a.npairs = npairs; a.coupled = coupled;
  a.P = c->geo.P.p; a.D = c->geo.D.p; a.kdop = c->d_kdop.p; a.steps = c->d_steps.p; a.offset = c->prm.offset;
  a.hit = c->self_hits.p; a.cap = cap; a.kmax = c->kmax.p;
  TOB_CUDA(c, cudaMemsetAsync(c->self_hits.p + cap, 0, 2 * sizeof(uint32_t), c->stream));
  if (n) {
    Prof prof(c, K_SELF_CCD);
    k_self_ccd_filter<<<div_up(n, 64), 64, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  k_self_ccd_resolve<<<1, 32, 0, c->stream>>>(a, steps_dev, c->self_hits.p + cap + 2);
  TOB_LAUNCH_CHECK(c);
  c->ctr.self_pairs += n;
  return 0;
}

// ---- front end: batched edge validity ------------------------------------------------------------------
// The reference's RRT motion validator (OMPL.cpp:36-96) checks one straight edge at a time: BVH::EdgeCollision (all points
// within d of the edge's box, BVH.cpp:95-133) and then CCD::GJKDCD(edge, point, d) (CCD.h:17-114) per candidate, invalid at
// the first collision.  Here many edges are checked in one launch: same walk as the broadphase with one "row" per edge, and
// every point that passes the leaf predicate runs GJK(2,1) in place; the answer per edge is an OR, so no candidate list.
struct EdgeArgs {
  BpArgs bp;
  const double* E;      // edges x 6: a xyz, b xyz
  double d;
  uint32_t* invalid;    // edges: set to 1 when some point is within d of the edge
};

__global__ void __launch_bounds__(BP_THREADS, 4) k_bp_edge(EdgeArgs a) {
  __shared__ BpShared s;
  uint32_t rank;
  const uint32_t n_items = bp_prepare(a.bp, s, &rank);
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double d2 = a.d * a.d;
  uint32_t h = 0;                                          // bp_item cursor
  for (uint32_t j = w; j < n_items; j += BP_WARPS) {
    uint32_t row, leaf;
    bp_item(a.bp, s, j, &h, &row, &leaf);
    double q[1][3];
    const bool ok = bp_point_test(a.bp, s, h, leaf * 32 + lane, &q[0][0], &q[0][1], &q[0][2]);
    if (!ok) continue;
    if (*((volatile uint32_t*)(a.invalid + row))) continue;     // already decided by another point
    double A[2][3], v[3];
    const double* e = a.E + (size_t)6 * row;
    for (int k = 0; k < 3; k++) { A[0][k] = e[k]; A[1][k] = e[3 + k]; }
    gjk_witness<2, 1>(A, q, v);
    const double dist2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (dist2 <= d2) a.invalid[row] = 1u;
  }
}

// edges_host: n x 2 x 3 (edge, endpoint, xyz).  valid_host[i] = 1 when no cloud point is within d of edge i.
int edge_validity(tob_ctx* c, const double* edges_host, int n, double d, uint8_t* valid_host) {
  if (c->n_pts == 0) return fail_msg(c, "tob_edge_validity: no point cloud uploaded");
  if (!c->cloud_n1.empty()) return fail_msg(c, "tob_edge_validity: one shared cloud is required (not per-problem clouds)");
  if (n <= 0) return 0;
  cudaStream_t st = c->stream;
  const uint32_t n1 = c->lvl[1].count;
  const int chunk_max = (int)std::min<uint64_t>((uint64_t)n, 0x7fffffffull / (n1 ? n1 : 1));
  if (chunk_max < 1) return fail_msg(c, "tob_edge_validity: cloud too large");
  std::vector<double> box((size_t)6 * n);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      const double p0 = edges_host[(size_t)6 * i + k], p1 = edges_host[(size_t)6 * i + 3 + k];
      // BVH.cpp:104-127: running min / max over the two endpoints, starting from +-inf
      double lo = INFINITY, hi = -INFINITY;
      if (p0 < lo) lo = p0; if (p0 > hi) hi = p0;
      if (p1 < lo) lo = p1; if (p1 > hi) hi = p1;
      box[(size_t)6 * i + k] = lo; box[(size_t)6 * i + 3 + k] = hi;
    }
  TOB_CUDA(c, c->scratch.ensure((size_t)12 * n));
  DBuf<uint32_t> inv;
  TOB_CUDA(c, inv.ensure(n));
  TOB_CUDA(c, cudaMemcpyAsync(c->scratch.p, box.data(), (size_t)6 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(c->scratch.p + (size_t)6 * n, edges_host, (size_t)6 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemsetAsync(inv.p, 0, (size_t)n * sizeof(uint32_t), st));
  for (int b0 = 0; b0 < n; b0 += chunk_max) {
    const int nb = std::min(chunk_max, n - b0);
    EdgeArgs a;
    a.bp.box = c->scratch.p + (size_t)6 * b0;
    a.bp.rows = (uint32_t)nb; a.bp.n1 = n1; a.bp.n_tasks = (uint32_t)nb * n1; a.bp.row_base = 0; a.bp.rows_all = (uint32_t)nb;
    a.bp.d = d; a.bp.row_task = nullptr; a.bp.row_l1 = nullptr; a.bp.cta_row = nullptr;
    for (int k = 0; k < 3; k++) {
      a.bp.l1lo[k] = c->lvl[1].lo[k]; a.bp.l1hi[k] = c->lvl[1].hi[k];
      a.bp.l0lo[k] = c->lvl[0].lo[k]; a.bp.l0hi[k] = c->lvl[0].hi[k];
    }
    a.bp.px = c->px.p; a.bp.py = c->py.p; a.bp.pz = c->pz.p;
    a.bp.bsum = nullptr; a.bp.cand_pt = nullptr; a.bp.cand_row = nullptr; a.bp.row_off = nullptr; a.bp.cand_cap = 0; a.bp.dc = c->dc.p;
    uint32_t tpc = a.bp.n_tasks / (4u * (uint32_t)c->sm_count);
    a.bp.tpc = tpc < 16u ? 16u : (tpc > (uint32_t)BP_THREADS ? (uint32_t)BP_THREADS : tpc);
    a.E = c->scratch.p + (size_t)6 * n + (size_t)6 * b0; a.d = d; a.invalid = inv.p + b0;
    k_bp_edge<<<div_up((size_t)a.bp.n_tasks, a.bp.tpc), BP_THREADS, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  std::vector<uint32_t> h(n);
  TOB_CUDA(c, cudaMemcpyAsync(h.data(), inv.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  for (int i = 0; i < n; i++) valid_host[i] = h[i] ? 0 : 1;
  return 0;
}

}  // namespace tob
