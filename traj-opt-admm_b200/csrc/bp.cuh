// bp.cuh -- block-cooperative walk of the 32-wide LBVH shared by the broadphase kernels (lbvh.cu: count / fill) and
// the fused CCD kernel (narrow.cu: swept-box traversal with the CCD ladder run in place on every hit point).
//
// Work decomposition.  One task = (row, level-1 node).  A CTA owns tpc <= BP_THREADS consecutive tasks:
//   phase A  thread per task: level-1 box test; hit tasks are compacted (ballot + prefix, deterministic) in shared memory
//   phase B  warp per hit task: 32 lanes test its 32 leaf boxes -> leaf mask
//   phase C  "items" = (hit task, hit leaf), enumerated task-major through a prefix over popc(leaf mask); the CTA's
//            warps take items round-robin, 32 lanes = the 32 points of the leaf.
// Items of one CTA are in (row, Morton position) order and CTAs cover consecutive tasks, so the candidate order is
// deterministic without atomics: position = (exclusive scan of the CTA totals) + (prefix inside the CTA).
// All warps of a CTA share its items, so one dense row (hundreds of hit leaves under a few level-1 nodes) no longer
// serialises on a single warp.
//
// The leaf predicate is the reference's (AABB.cc:131-161 as called at :647), FP64, unfused:
//     candidate  <=>  for every axis:  !(p + d < lo)  &&  !(p > hi + d)
#pragma once
#include "ctx.cuh"

namespace tob {

#define BP_THREADS 128
#define BP_WARPS (BP_THREADS / 32)
#define BP_BATCH 512   // items staged per batch by the fill pass

struct BpArgs {
  const double* box;           // rows_all x 6 (lo xyz, hi xyz)
  uint32_t rows, n1, n_tasks, row_base, rows_all;
  uint32_t tpc;                // tasks per CTA (<= BP_THREADS): small queries are spread over more CTAs
  double d;
  const double *l1lo[3], *l1hi[3], *l0lo[3], *l0hi[3];
  const double *px, *py, *pz;
  uint32_t* bsum;              // per-CTA candidate totals (count: out; fill: exclusive scan, in)
  uint32_t *cand_pt, *cand_row, *row_off;
  uint32_t cand_cap;
  DevCounts* dc;
  // one cloud per robot (batched independent problems): row r walks level-1 nodes [row_l1[r], row_l1[r] + n1(r)) and its
  // tasks are [row_task[r], row_task[r+1]); null = one shared cloud, n1 nodes for every row
  const uint32_t *row_task, *row_l1;
  const uint32_t* cta_row;     // per-robot clouds, whole-context query: row of the first task of every CTA (host-built), or null
  // item records count -> fill (rec_cap = 0: none, the fill pass repeats the tests)
  uint32_t *rec_row, *rec_leaf, *rec_pm, *cta_ibase, *cta_nitems, *row_li;
  uint32_t rec_cap;
};

struct BpShared {
  uint32_t hit_task[BP_THREADS];      // local task index of the h-th hit task
  uint32_t hit_row[BP_THREADS];       // its global row
  uint32_t hit_nd[BP_THREADS];        // its level-1 node (global index)
  double hit_q[BP_THREADS][6];        // the query box of its row (lo xyz, hi xyz): read once from global memory in phase A
  uint32_t lmask[BP_THREADS];         // its leaf mask
  uint32_t item_base[BP_THREADS + 1]; // exclusive prefix of popc(lmask)
  uint32_t wtmp[BP_WARPS + 1];
  uint32_t n_hit;
};

void bp_args(tob_ctx* c, int row_base, int rows, double d, BpArgs& a);   // lbvh.cu

// the reference predicate with the query box [qlo,qhi] as "this" and the node/point as the argument
__device__ __forceinline__ bool box_hit(double nlo, double nhi, double qlo, double qhi, double d) {
  return !(nhi + d < qlo) && !(nlo > qhi + d);
}
// all three axes, operands already in registers.  The callers load every operand BEFORE the tests: written as
// box_hit(load x) && box_hit(load y) && ... the short-circuit evaluation turns the loads of y and z into dependent global
// loads behind a branch, three memory round trips instead of one (measured: 3.4 k cycles per phase, 2.6 k per item).
__device__ __forceinline__ bool box_hit3(const double* nlo, const double* nhi, const double* q, double d) {
  const bool hx = box_hit(nlo[0], nhi[0], q[0], q[3], d), hy = box_hit(nlo[1], nhi[1], q[1], q[4], d),
             hz = box_hit(nlo[2], nhi[2], q[2], q[5], d);
  return hx & hy & hz;
}

// exclusive prefix of v over the CTA (BP_THREADS threads); *total = CTA sum.  Contains two __syncthreads().
__device__ __forceinline__ uint32_t bp_block_excl(uint32_t v, uint32_t* wtmp, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                       // wtmp may still be read from a previous call
  if (lane == 31) wtmp[w] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < BP_WARPS; i++) {
    uint32_t x = wtmp[i];
    if (i < (int)w) base += x;
    tot += x;
  }
  *total = tot;
  return base + inc - v;
}

// One CTA of 1024 threads: exclusive scan of v[0..n) in place; returns the total (uniform).  Used for the short arrays of
// per-CTA / per-chunk / per-row totals that sit between a count pass and a fill pass.  Four consecutive elements per thread
// and tile (the array of per-chunk plane counts of the 1024-problem batch has 573 k entries: 140 tiles instead of 560).
__device__ __forceinline__ uint32_t cta1024_scan_inplace(uint32_t* v, uint32_t n) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t s_tile;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += 4096) {
    const uint32_t k = base + 4 * threadIdx.x;
    uint32_t x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = k + i < n ? v[k + i] : 0;
    const uint32_t mine = (x[0] + x[1]) + (x[2] + x[3]);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint32_t sv = wsum[lane];
      uint32_t si = sv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - sv;
      if (lane == 31) s_tile = si;
    }
    __syncthreads();
    uint32_t e = carry + wsum[w] + inc - mine;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (k + i < n) v[k + i] = e;
      e += x[i];
    }
    carry += s_tile;
    __syncthreads();
  }
  return carry;
}

// task (relative to the first queried row) -> global row, global level-1 node, "first node of its row".
// row_hint: a row known to be <= the task's row (per-robot clouds: the row of the CTA's first task, found once per CTA by
// binary search; the tasks of a CTA are consecutive, so every thread then walks a few rows forward instead of searching)
__device__ __forceinline__ void bp_task(const BpArgs& a, uint32_t t, uint32_t* row, uint32_t* nd, bool* first, uint32_t row_hint) {
  if (a.row_task == nullptr) {
    const uint32_t r = t / a.n1;
    *row = a.row_base + r;
    *nd = t - r * a.n1;
    *first = t == r * a.n1;
    return;
  }
  const uint32_t tg = t + a.row_task[a.row_base];
  uint32_t lo = row_hint;
  const uint32_t last = a.row_base + a.rows - 1;
  while (lo < last && a.row_task[lo + 1] <= tg) lo++;      // largest row with row_task[row] <= tg
  *row = lo;
  *nd = a.row_l1[lo] + (tg - a.row_task[lo]);
  *first = tg == a.row_task[lo];
}

// row of the first task of this CTA (per-robot clouds): from the host-built table, else by binary search (16 dependent
// loads on the critical path of every CTA of a 65536-row batch); a.row_base when all rows share one cloud
__device__ __forceinline__ uint32_t bp_first_row(const BpArgs& a) {
  if (a.row_task == nullptr) return a.row_base;
  if (a.cta_row != nullptr) return a.cta_row[blockIdx.x];
  const uint32_t tg = blockIdx.x * a.tpc + a.row_task[a.row_base];
  uint32_t lo = a.row_base, hi = a.row_base + a.rows;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.row_task[mid] <= tg) lo = mid; else hi = mid;
  }
  return lo;
}

// phases A + B + item prefix.  Returns the number of items of this CTA; *my_rank = hit tasks before this thread's task.
// *my_row / *my_first (optional): the row of this thread's task and whether it is the first task of its row.
__device__ __forceinline__ uint32_t bp_prepare(const BpArgs& a, BpShared& s, uint32_t* my_rank, uint32_t* my_row = nullptr,
                                              bool* my_first = nullptr) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t t = blockIdx.x * a.tpc + tid;
  bool hit = false;
  uint32_t row = 0, nd = 0;
  double q[6] = {0, 0, 0, 0, 0, 0};
  if (my_first) *my_first = false;
  if (tid < a.tpc && t < a.n_tasks) {
    bool first;
    bp_task(a, t, &row, &nd, &first, bp_first_row(a));
    if (my_row) *my_row = row;
    if (my_first) *my_first = first;
    const double* qb = a.box + (size_t)6 * row;
    double nlo[3], nhi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { nlo[k] = a.l1lo[k][nd]; nhi[k] = a.l1hi[k][nd]; }
#pragma unroll
    for (int k = 0; k < 6; k++) q[k] = qb[k];
    hit = box_hit3(nlo, nhi, q, a.d);
  }
  uint32_t n_hit;
  const uint32_t rank = bp_block_excl(hit ? 1u : 0u, s.wtmp, &n_hit);
  *my_rank = rank;
  if (hit) {
    s.hit_task[rank] = tid; s.hit_row[rank] = row; s.hit_nd[rank] = nd;
#pragma unroll
    for (int k = 0; k < 6; k++) s.hit_q[rank][k] = q[k];
  }
  if (tid == 0) s.n_hit = n_hit;
  __syncthreads();
  for (uint32_t h = w; h < n_hit; h += BP_WARPS) {
    const uint32_t leaf = s.hit_nd[h] * 32 + lane;   // level-0 arrays are padded to 32 with empty boxes
    double nlo[3], nhi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { nlo[k] = a.l0lo[k][leaf]; nhi[k] = a.l0hi[k][leaf]; }
    const bool lh = box_hit3(nlo, nhi, s.hit_q[h], a.d);
    const uint32_t lm = __ballot_sync(0xffffffffu, lh);
    if (lane == 0) s.lmask[h] = lm;
  }
  __syncthreads();
  uint32_t n_items;
  const uint32_t ib = bp_block_excl(tid < n_hit ? (uint32_t)__popc(s.lmask[tid]) : 0u, s.wtmp, &n_items);
  if (tid < n_hit) s.item_base[tid] = ib;
  if (tid == 0) s.item_base[n_hit] = n_items;
  __syncthreads();
  return n_items;
}

// item j -> hit task h and the leaf it addresses (called by all 32 lanes of a warp with the same j).  *h_out is a cursor: a warp visits its items in increasing order, so the
// search walks forward from the task of the previous item (start with 0); uniform in the warp
__device__ __forceinline__ void bp_item(const BpArgs& a, const BpShared& s, uint32_t j, uint32_t* h_out, uint32_t* row, uint32_t* leaf) {
  uint32_t lo = *h_out;                     // largest h with item_base[h] <= j
  while (lo + 1 < s.n_hit && s.item_base[lo + 1] <= j) lo++;
  const uint32_t k = j - s.item_base[lo];
  // position of the k-th set bit of the (warp-uniform) leaf mask: lane l owns bit l, the lane whose bit is set with k set bits
  // below it answers (__fns is emulated in software: it was 20 % of the stall samples of k_bp_count, profiles/r02z_hot_*)
  const uint32_t m = s.lmask[lo], lane = threadIdx.x & 31;
  const uint32_t lb = (uint32_t)__ffs(__ballot_sync(0xffffffffu, ((m >> lane) & 1u) && (uint32_t)__popc(m & ((1u << lane) - 1u)) == k)) - 1u;
  *h_out = lo;
  *row = s.hit_row[lo];
  *leaf = s.hit_nd[lo] * 32 + lb;
}

// 32 lanes = the 32 points of a leaf against the box of hit task h (its row's box sits in shared memory)
__device__ __forceinline__ bool bp_point_test(const BpArgs& a, const BpShared& s, uint32_t h, uint32_t p, double* x, double* y, double* z) {
  const double pt[3] = {a.px[p], a.py[p], a.pz[p]};
  *x = pt[0]; *y = pt[1]; *z = pt[2];
  return box_hit3(pt, pt, s.hit_q[h], a.d);
}

}  // namespace tob
