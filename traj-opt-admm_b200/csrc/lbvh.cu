// lbvh.cu -- GPU LBVH over the obstacle cloud and the bit-exact broadphase.
//
// Replaces: BVH::InitPointcloud (HighOrderCCD/BVH/BVH.cpp:53-92, incremental SAH tree of aabb::Tree) and the
// tree queries behind BVH::DCDCollision / CCDCollision (BVH.cpp:149-249 -> aabb::Tree::query AABB.cc:608-667).
// Only the LEAF PREDICATE of the reference is contractual (AABB.cc:131-161 as called at :647):
//     candidate  <=>  for every axis:  !(p + d < lo)  &&  !(p > hi + d)         (FP64, unfused)
// Node boxes here are exact FP64 min/max of their points, so testing them with the same expression can never
// reject a subtree that holds a candidate (rounding is monotone): the candidate SET equals the reference's.
//
// Layout in HBM (all SoA, FP64):
//   px,py,pz[n_pad], pid[n_pad]   points in Morton order, padded to a multiple of 32 with +inf
//   level 0: one box per 32 consecutive points (a leaf = one warp-wide coalesced load)
//   level 1: one box per 32 leaves.  Boxes are lo[3][count_pad], hi[3][count_pad].
// Broadphase work decomposition: one task = (row, level-1 node); a CTA owns 128 consecutive tasks and shares the hit
// leaves among its warps (bp.cuh).  Two passes (count per CTA, scan of the CTA totals, fill) give a deterministic
// candidate order (row, Morton position) without atomics; sizes stay on the device (DevCounts), nothing is read back.
//
// Algorithmic bytes: build 128 B/point; query 48 B/row + 16 B/(row x L1 node) + 28 B/candidate.
#include <string.h>

#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"
#include "bp.cuh"

namespace tob {

// ---- build -------------------------------------------------------------------------------------------------
__global__ void k_minmax(const double* __restrict__ V, uint32_t n, double* __restrict__ part) {
  // V column-major n x 3; part: gridDim.x x 6
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int a = 0; a < 3; a++) {
      double v = V[(size_t)a * n + i];
      lo[a] = fmin(lo[a], v);
      hi[a] = fmax(hi[a], v);
    }
  __shared__ double sm[6][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { sm[a][w] = lo[a]; sm[3 + a][w] = hi[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int nw = blockDim.x >> 5;
    double r = sm[threadIdx.x][0];
    for (int i = 1; i < nw; i++) r = threadIdx.x < 3 ? fmin(r, sm[threadIdx.x][i]) : fmax(r, sm[threadIdx.x][i]);
    part[blockIdx.x * 6 + threadIdx.x] = r;
  }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

// second stage of the cloud bounds: folds the per-CTA partials; bnd = lo[3] | scale[3] (quantisation to 21 bits per axis)
__global__ void k_minmax_final(const double* __restrict__ part, int nb, double* __restrict__ bnd) {
  const int a = threadIdx.x;
  if (a >= 3) return;
  double lo = INFINITY, hi = -INFINITY;
  for (int b = 0; b < nb; b++) {
    lo = fmin(lo, part[b * 6 + a]);
    hi = fmax(hi, part[b * 6 + 3 + a]);
  }
  const double ext = hi - lo;
  bnd[a] = lo;
  bnd[3 + a] = ext > 0 ? 2097151.0 / ext : 0.0;
}

__global__ void k_morton(const double* __restrict__ V, uint32_t n, const double* __restrict__ bnd, uint64_t* __restrict__ key,
                         uint32_t* __restrict__ idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lx = bnd[0], ly = bnd[1], lz = bnd[2], sx = bnd[3], sy = bnd[4], sz = bnd[5];
  double x = (V[i] - lx) * sx, y = (V[(size_t)n + i] - ly) * sy, z = (V[(size_t)2 * n + i] - lz) * sz;
  uint64_t qx = (uint64_t)fmin(fmax(x, 0.0), 2097151.0);
  uint64_t qy = (uint64_t)fmin(fmax(y, 0.0), 2097151.0);
  uint64_t qz = (uint64_t)fmin(fmax(z, 0.0), 2097151.0);
  key[i] = spread21(qx) | (spread21(qy) << 1) | (spread21(qz) << 2);
  idx[i] = i;
}

__global__ void k_gather(const double* __restrict__ V, uint32_t n, uint32_t n_pad, const uint32_t* __restrict__ idx,
                         double* __restrict__ px, double* __restrict__ py, double* __restrict__ pz, uint32_t* __restrict__ pid) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i < n) {
    uint32_t s = idx[i];
    px[i] = V[s]; py[i] = V[(size_t)n + s]; pz[i] = V[(size_t)2 * n + s];
    pid[i] = s;
  } else {
    px[i] = INFINITY; py[i] = INFINITY; pz[i] = INFINITY;   // never a candidate: p > hi + d
    pid[i] = 0xffffffffu;
  }
}

// one warp per parent: box of its 32 children (children are points when from_points)
__global__ void k_level(int from_points, uint32_t n_child, uint32_t n_parent_pad, const double* __restrict__ cx_lo,
                        const double* __restrict__ cy_lo, const double* __restrict__ cz_lo, const double* __restrict__ cx_hi,
                        const double* __restrict__ cy_hi, const double* __restrict__ cz_hi, double* __restrict__ ox_lo,
                        double* __restrict__ oy_lo, double* __restrict__ oz_lo, double* __restrict__ ox_hi,
                        double* __restrict__ oy_hi, double* __restrict__ oz_hi) {
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_parent_pad) return;
  uint32_t ch = w * 32 + lane;
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (ch < n_child) {
    lo[0] = cx_lo[ch]; lo[1] = cy_lo[ch]; lo[2] = cz_lo[ch];
    hi[0] = cx_hi[ch]; hi[1] = cy_hi[ch]; hi[2] = cz_hi[ch];
    if (from_points && isinf(lo[0])) { lo[0] = lo[1] = lo[2] = INFINITY; hi[0] = hi[1] = hi[2] = -INFINITY; }
  }
  for (int a = 0; a < 3; a++)
    for (int o = 16; o; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  if (lane == 0) {
    ox_lo[w] = lo[0]; oy_lo[w] = lo[1]; oz_lo[w] = lo[2];
    ox_hi[w] = hi[0]; oy_hi[w] = hi[1]; oz_hi[w] = hi[2];
  }
}

// Morton-sort the clouds into the concatenated SoA arrays: cloud b goes to point offset base[b] and owns slot[b] points
// (padding = +inf points that can never be candidates).  One workspace sized for the largest cloud serves all of them and
// nothing is read back per cloud: the bounds stay on the device, the host only refills one of two pinned staging buffers
// while the stream sorts the previous cloud.
static int lbvh_sort_clouds(tob_ctx* c, const double* const* V_host, const uint32_t* n, const size_t* base, const uint32_t* slot,
                            int n_clouds) {
  cudaStream_t st = c->stream;
  uint32_t nmax = 0;
  for (int b = 0; b < n_clouds; b++) nmax = n[b] > nmax ? n[b] : nmax;
  const int NB = 296;
  DBuf<double> V, part, bnd;
  DBuf<uint64_t> key, key2;
  DBuf<uint32_t> idx, idx2;
  DBuf<uint8_t> tmp;
  TOB_CUDA(c, V.ensure((size_t)3 * nmax)); TOB_CUDA(c, part.ensure(NB * 6)); TOB_CUDA(c, bnd.ensure(8));
  TOB_CUDA(c, key.ensure(nmax)); TOB_CUDA(c, key2.ensure(nmax)); TOB_CUDA(c, idx.ensure(nmax)); TOB_CUDA(c, idx2.ensure(nmax));
  size_t tmp_bytes = 0;
  TOB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, idx.p, idx2.p, (int)nmax, 0, 63, st));
  TOB_CUDA(c, tmp.ensure(tmp_bytes));
  struct Pinned {
    double* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~Pinned() { for (int i = 0; i < 2; i++) { if (p[i]) cudaFreeHost(p[i]); if (ev[i]) cudaEventDestroy(ev[i]); } }
  } pin;
  for (int i = 0; i < 2; i++) {
    TOB_CUDA(c, cudaMallocHost((void**)&pin.p[i], (size_t)3 * nmax * sizeof(double)));
    TOB_CUDA(c, cudaEventCreateWithFlags(&pin.ev[i], cudaEventDisableTiming));
  }
  cudaEvent_t t0, t1;
  TOB_CUDA(c, cudaEventCreate(&t0)); TOB_CUDA(c, cudaEventCreate(&t1));
  TOB_CUDA(c, cudaEventRecord(t0, st));
  for (int b = 0; b < n_clouds; b++) {
    const int q = b & 1;
    const uint32_t nb = n[b];
    if (b >= 2) TOB_CUDA(c, cudaEventSynchronize(pin.ev[q]));          // the copy out of this staging buffer has finished
    memcpy(pin.p[q], V_host[b], (size_t)3 * nb * sizeof(double));
    TOB_CUDA(c, cudaMemcpyAsync(V.p, pin.p[q], (size_t)3 * nb * sizeof(double), cudaMemcpyHostToDevice, st));
    TOB_CUDA(c, cudaEventRecord(pin.ev[q], st));
    const int blocks = nb < 65536 ? 32 : NB;
    k_minmax<<<blocks, 256, 0, st>>>(V.p, nb, part.p);
    TOB_LAUNCH_CHECK(c);
    k_minmax_final<<<1, 32, 0, st>>>(part.p, blocks, bnd.p);
    TOB_LAUNCH_CHECK(c);
    k_morton<<<div_up(nb, 256), 256, 0, st>>>(V.p, nb, bnd.p, key.p, idx.p);
    TOB_LAUNCH_CHECK(c);
    TOB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, idx.p, idx2.p, (int)nb, 0, 63, st));
    k_gather<<<div_up(slot[b], 256), 256, 0, st>>>(V.p, nb, slot[b], idx2.p, c->px.p + base[b], c->py.p + base[b], c->pz.p + base[b],
                                                  c->pid.p + base[b]);
    TOB_LAUNCH_CHECK(c);
  }
  TOB_CUDA(c, cudaEventRecord(t1, st));
  cudaError_t e = cudaStreamSynchronize(st);
  float ms = 0;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0); cudaEventDestroy(t1);
  TOB_CUDA(c, e);
  c->build_ms = ms;
  return 0;
}

// levels over the concatenated, padded point array: 0 = leaves (32 points), then 32-ary up to a single node
static int lbvh_levels(tob_ctx* c, size_t n_pad) {
  cudaStream_t st = c->stream;
  uint32_t counts[TOB_MAX_LEVELS], pads[TOB_MAX_LEVELS];
  int nl = 0;
  uint32_t cnt = (uint32_t)(n_pad / 32);
  while (true) {
    counts[nl] = cnt;
    pads[nl] = (cnt + 31u) & ~31u;
    nl++;
    if (cnt == 1 || nl == TOB_MAX_LEVELS) break;
    cnt = (cnt + 31) / 32;
  }
  if (nl < 2) { counts[1] = 1; pads[1] = 32; nl = 2; }   // always have a level 1 (task level of the broadphase)
  size_t tot = 0;
  for (int l = 0; l < nl; l++) tot += (size_t)6 * pads[l];
  TOB_CUDA(c, c->lvl_store.ensure(tot));
  size_t off = 0;
  for (int l = 0; l < nl; l++) {
    c->lvl[l].count = counts[l];
    for (int a = 0; a < 3; a++) { c->lvl[l].lo[a] = c->lvl_store.p + off; off += pads[l]; }
    for (int a = 0; a < 3; a++) { c->lvl[l].hi[a] = c->lvl_store.p + off; off += pads[l]; }
  }
  c->n_levels = nl;
  for (int l = 0; l < nl; l++) {
    Level& L = c->lvl[l];
    uint32_t n_child = (l == 0) ? (uint32_t)n_pad : counts[l - 1];
    const double *a0, *a1, *a2, *b0, *b1, *b2;
    if (l == 0) { a0 = b0 = c->px.p; a1 = b1 = c->py.p; a2 = b2 = c->pz.p; }
    else { Level& C = c->lvl[l - 1]; a0 = C.lo[0]; a1 = C.lo[1]; a2 = C.lo[2]; b0 = C.hi[0]; b1 = C.hi[1]; b2 = C.hi[2]; }
    k_level<<<div_up((size_t)pads[l] * 32, 256), 256, 0, st>>>(l == 0, n_child, pads[l], a0, a1, a2, b0, b1, b2, L.lo[0], L.lo[1],
                                                                  L.lo[2], L.hi[0], L.hi[1], L.hi[2]);
    TOB_LAUNCH_CHECK(c);
  }
  c->h_pid.resize(n_pad);
  TOB_CUDA(c, cudaMemcpyAsync(c->h_pid.data(), c->pid.p, n_pad * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  return 0;
}

// one cloud shared by every robot of the context (BVH::InitPointcloud)
int lbvh_build(tob_ctx* c, const double* V_host, uint32_t n) {
  if (n == 0) return fail_msg(c, "tob_cloud_upload: empty cloud");
  uint32_t n_pad = (n + 31u) & ~31u;
  TOB_CUDA(c, c->px.ensure(n_pad)); TOB_CUDA(c, c->py.ensure(n_pad)); TOB_CUDA(c, c->pz.ensure(n_pad));
  TOB_CUDA(c, c->pid.ensure(n_pad));
  const size_t base0 = 0;
  TOB_TRY(lbvh_sort_clouds(c, &V_host, &n, &base0, &n_pad, 1));
  TOB_TRY(lbvh_levels(c, n_pad));
  c->n_pts = n; c->n_pad = n_pad; c->build_points = n;
  c->cloud_n1.clear(); c->cloud_l1.clear();
  c->row_task.release(); c->row_l1.release();
  return 0;
}

// one cloud per robot slot (batched independent problems): clouds are concatenated, each padded to a multiple of 1024
// points so that neither a leaf (32 points) nor a level-1 node (32 leaves) straddles two clouds.  Per row the broadphase
// then walks only the level-1 nodes of its own cloud (row_task / row_l1).
static uint32_t bp_tpc(tob_ctx* c, uint32_t n_tasks);

int lbvh_build_batch(tob_ctx* c, const double* const* V_host, const uint32_t* n, int n_clouds) {
  if (n_clouds != c->n_robots()) return fail_msg(c, "tob_cloud_upload_batch: one cloud per robot slot (uav_num) is required");
  size_t total = 0;
  std::vector<size_t> base(n_clouds);
  std::vector<uint32_t> slot(n_clouds);
  for (int b = 0; b < n_clouds; b++) {
    if (n[b] == 0) return fail_msg(c, "tob_cloud_upload_batch: empty cloud");
    base[b] = total;
    slot[b] = (n[b] + 1023u) & ~1023u;
    total += slot[b];
  }
  if (total > 0xfff00000ull) return fail_msg(c, "tob_cloud_upload_batch: more than 2^32 points");
  TOB_CUDA(c, c->px.ensure(total)); TOB_CUDA(c, c->py.ensure(total)); TOB_CUDA(c, c->pz.ensure(total));
  TOB_CUDA(c, c->pid.ensure(total));
  TOB_TRY(lbvh_sort_clouds(c, V_host, n, base.data(), slot.data(), n_clouds));
  TOB_TRY(lbvh_levels(c, total));
  c->build_points = 0;
  for (int b = 0; b < n_clouds; b++) c->build_points += n[b];
  c->n_pts = (uint32_t)total; c->n_pad = (uint32_t)total;
  c->cloud_n1.resize(n_clouds); c->cloud_l1.resize(n_clouds);
  for (int b = 0; b < n_clouds; b++) { c->cloud_l1[b] = (uint32_t)(base[b] / 1024); c->cloud_n1[b] = slot[b] / 1024; }
  // per-row task prefix and level-1 base
  const int rows = c->rows_all();
  std::vector<uint32_t> rt(rows + 1), rl(rows + 1);
  uint64_t acc = 0;
  for (int r = 0; r < rows; r++) {
    const int u = r / c->n_tr;
    rt[r] = (uint32_t)acc; rl[r] = c->cloud_l1[u];
    acc += c->cloud_n1[u];
    if (acc > 0xfff00000ull) return fail_msg(c, "tob_cloud_upload_batch: more than 2^32 broadphase tasks");
  }
  rt[rows] = (uint32_t)acc; rl[rows] = 0;
  TOB_CUDA(c, c->row_task.ensure(rows + 1)); TOB_CUDA(c, c->row_l1.ensure(rows + 1));
  TOB_CUDA(c, cudaMemcpyAsync(c->row_task.p, rt.data(), (rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  TOB_CUDA(c, cudaMemcpyAsync(c->row_l1.p, rl.data(), (rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  // row of the first task of every CTA of a whole-context query (bp.cuh: bp_first_row)
  {
    const uint32_t n_tasks = (uint32_t)acc, tpc = bp_tpc(c, n_tasks), nblk = (n_tasks + tpc - 1) / tpc;
    std::vector<uint32_t> cr(nblk + 1);
    uint32_t r = 0;
    for (uint32_t b = 0; b < nblk; b++) {
      const uint32_t tg = b * tpc;
      while (r + 1 < (uint32_t)rows && rt[r + 1] <= tg) r++;
      cr[b] = r;
    }
    TOB_CUDA(c, c->cta_row.ensure(nblk + 1));
    TOB_CUDA(c, cudaMemcpyAsync(c->cta_row.p, cr.data(), nblk * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    c->cta_row_tpc = tpc;
  }
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->h_row_task = rt;
  return 0;
}

// ---- broadphase ------------------------------------------------------------------------------------------------
// count pass: candidates per CTA (bp.cuh explains the walk)
__global__ void __launch_bounds__(BP_THREADS) k_bp_count(BpArgs a) {
  __shared__ BpShared s;
  __shared__ uint32_t s_ibase;
  uint32_t rank, my_row;
  bool my_first;
  const uint32_t n_items = bp_prepare(a, s, &rank, &my_row, &my_first);
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // room for this CTA's item records (row, leaf, point mask): the fill pass then scatters without walking the tree again.
  // The order in which CTAs get their room is arbitrary and irrelevant (candidate positions come from the scanned totals).
  if (threadIdx.x == 0) {
    uint32_t ib = 0xffffffffu;
    if (a.rec_cap && n_items) {
      ib = atomicAdd(&a.dc->bp_items, n_items);
      if (ib > a.rec_cap || n_items > a.rec_cap - ib) ib = 0xffffffffu;     // no room: the fill pass repeats the tests for this CTA
    } else if (a.rec_cap) ib = 0;
    s_ibase = ib;
    a.cta_ibase[blockIdx.x] = ib;
    a.cta_nitems[blockIdx.x] = n_items;
  }
  if (my_first && a.rec_cap) a.row_li[my_row] = s.item_base[rank];          // items of this CTA in front of the row's first task
  __syncthreads();
  const uint32_t ibase = s_ibase;
  uint32_t cnt = 0;
  uint32_t h = 0;                                          // bp_item cursor
  // two items of the warp in flight: the point loads of the second are issued before the first is tested (the kernel waits
  // on these loads: long scoreboard, 5.2 cycles per issue)
  for (uint32_t j = w; j < n_items; j += 2 * BP_WARPS) {
    uint32_t row0, leaf0, row1 = 0, leaf1 = 0;
    bp_item(a, s, j, &h, &row0, &leaf0);
    const uint32_t h0 = h;
    const bool two = j + BP_WARPS < n_items;               // uniform in the warp
    if (two) bp_item(a, s, j + BP_WARPS, &h, &row1, &leaf1);
    const uint32_t p0 = leaf0 * 32 + lane, p1 = (two ? leaf1 : leaf0) * 32 + lane;
    const double pa[3] = {a.px[p0], a.py[p0], a.pz[p0]}, pb[3] = {a.px[p1], a.py[p1], a.pz[p1]};
    const bool ok0 = box_hit3(pa, pa, s.hit_q[h0], a.d), ok1 = two && box_hit3(pb, pb, s.hit_q[h], a.d);
    const uint32_t pm0 = __ballot_sync(0xffffffffu, ok0), pm1 = __ballot_sync(0xffffffffu, ok1);
    cnt += __popc(pm0) + __popc(pm1);
    if (ibase != 0xffffffffu && lane == 0) {
      a.rec_row[ibase + j] = row0; a.rec_leaf[ibase + j] = leaf0; a.rec_pm[ibase + j] = pm0;
      if (two) { a.rec_row[ibase + j + BP_WARPS] = row1; a.rec_leaf[ibase + j + BP_WARPS] = leaf1; a.rec_pm[ibase + j + BP_WARPS] = pm1; }
    }
  }
  __syncthreads();
  if (lane == 0) s.wtmp[w] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int i = 0; i < BP_WARPS; i++) tot += s.wtmp[i];
    a.bsum[blockIdx.x] = tot;
  }
}

// one CTA: exclusive scan of the per-CTA totals; total -> dc->n_cand, sticky overflow bit, cumulative counter
__global__ void __launch_bounds__(1024) k_bp_top(uint32_t* bsum, uint32_t nblk, uint32_t cap, int count_as, DevCounts* dc) {
  const uint32_t total = cta1024_scan_inplace(bsum, nblk);
  if (threadIdx.x == 0) {
    bsum[nblk] = total;
    dc->n_cand = total;
    if (total > cap) dc->overflow |= TOB_OVF_CAND;
    else if (count_as == 1) dc->dcd_candidates += total;
  }
}

// fill pass: same walk; items are staged in batches so that the position of every candidate is
//   (scanned CTA base) + (candidates of earlier items of this CTA) + (rank among the lanes of its leaf)
__global__ void __launch_bounds__(BP_THREADS) k_bp_fill(BpArgs a) {
  __shared__ BpShared s;
  __shared__ uint32_t s_pm[BP_BATCH], s_pre[BP_BATCH], s_leaf[BP_BATCH], s_row[BP_BATCH];
  __shared__ uint32_t s_taskcand[BP_THREADS + 1];   // candidates of this CTA before hit task h
  const uint32_t total = a.dc->n_cand;
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // candidate offsets of the rows outside the queried range (empty lists)
  if (blockIdx.x == 0)
    for (uint32_t g = tid; g <= a.rows_all; g += BP_THREADS)
      if (g < a.row_base || g >= a.row_base + a.rows) a.row_off[g] = g < a.row_base ? 0u : total;
  if (blockIdx.x == 0 && tid == 0) a.dc->bp_items = 0;     // the records of this query are handed out: room for the next count pass
  if (total > a.cand_cap) return;   // overflow: the host grows the buffers and runs the query again
  const uint32_t ibase = a.rec_cap ? a.cta_ibase[blockIdx.x] : 0xffffffffu;
  if (ibase != 0xffffffffu) {
    // the count pass left (row, leaf, point mask) of every item of this CTA: prefix over the masks and scatter
    const uint32_t n_it = a.cta_nitems[blockIdx.x], cta_base = a.bsum[blockIdx.x];
    const uint32_t t = blockIdx.x * a.tpc + tid;
    uint32_t row = 0, nd, li = 0xffffffffu, off = 0;
    bool first = false;
    if (tid < a.tpc && t < a.n_tasks) {
      bp_task(a, t, &row, &nd, &first, bp_first_row(a));
      if (first) li = a.row_li[row];
    }
    uint32_t run = 0;
    for (uint32_t b0 = 0; b0 < n_it; b0 += BP_BATCH) {
      const uint32_t nb = min((uint32_t)BP_BATCH, n_it - b0);
      constexpr int PER = BP_BATCH / BP_THREADS;
      uint32_t loc[PER], sum = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) {                          // PER consecutive items per thread
        const uint32_t jj = tid * PER + i;
        uint32_t pm = 0;
        if (jj < nb) {
          pm = a.rec_pm[ibase + b0 + jj];
          s_pm[jj] = pm; s_leaf[jj] = a.rec_leaf[ibase + b0 + jj]; s_row[jj] = a.rec_row[ibase + b0 + jj];
        }
        loc[i] = (uint32_t)__popc(pm);
        sum += loc[i];
      }
      uint32_t btot;
      uint32_t e = bp_block_excl(sum, s.wtmp, &btot);          // two barriers: s_pm / s_leaf / s_row are visible afterwards
#pragma unroll
      for (int i = 0; i < PER; i++) {
        const uint32_t jj = tid * PER + i;
        if (jj < nb) s_pre[jj] = e;
        e += loc[i];
      }
      __syncthreads();
      if (first && li >= b0 && li < b0 + nb) off = run + s_pre[li - b0];
      for (uint32_t jj = w; jj < nb; jj += BP_WARPS) {
        const uint32_t pm = s_pm[jj];
        if (pm & (1u << lane)) {
          const uint32_t pos = cta_base + run + s_pre[jj] + __popc(pm & ((1u << lane) - 1u));
          a.cand_pt[pos] = s_leaf[jj] * 32 + lane;
          a.cand_row[pos] = s_row[jj];
        }
      }
      run += btot;
      __syncthreads();
    }
    if (first) a.row_off[row] = cta_base + (li >= n_it ? run : off);
    return;
  }
  uint32_t rank;
  const uint32_t n_items = bp_prepare(a, s, &rank);
  const uint32_t n_hit = s.n_hit;
  const uint32_t cta_base = a.bsum[blockIdx.x];
  uint32_t run = 0;
  uint32_t h = 0;                                          // bp_item cursor (a warp's items increase across batches too)
  for (uint32_t b0 = 0; b0 < n_items; b0 += BP_BATCH) {
    const uint32_t nb = min((uint32_t)BP_BATCH, n_items - b0);
    for (uint32_t jj = w; jj < nb; jj += 2 * BP_WARPS) {     // two items in flight, as in k_bp_count
      uint32_t row0, leaf0, row1 = 0, leaf1 = 0;
      bp_item(a, s, b0 + jj, &h, &row0, &leaf0);
      const uint32_t h0 = h;
      const bool two = jj + BP_WARPS < nb;                   // uniform in the warp
      if (two) bp_item(a, s, b0 + jj + BP_WARPS, &h, &row1, &leaf1);
      const uint32_t p0 = leaf0 * 32 + lane, p1 = (two ? leaf1 : leaf0) * 32 + lane;
      const double pa[3] = {a.px[p0], a.py[p0], a.pz[p0]}, pb[3] = {a.px[p1], a.py[p1], a.pz[p1]};
      const uint32_t pm0 = __ballot_sync(0xffffffffu, box_hit3(pa, pa, s.hit_q[h0], a.d));
      const uint32_t pm1 = __ballot_sync(0xffffffffu, two && box_hit3(pb, pb, s.hit_q[h], a.d));
      if (lane == 0) {
        s_pm[jj] = pm0; s_leaf[jj] = leaf0; s_row[jj] = row0;
        if (two) { s_pm[jj + BP_WARPS] = pm1; s_leaf[jj + BP_WARPS] = leaf1; s_row[jj + BP_WARPS] = row1; }
      }
    }
    __syncthreads();
    // exclusive prefix of popc(pm) over the batch: BP_BATCH / BP_THREADS consecutive items per thread
    constexpr int PER = BP_BATCH / BP_THREADS;
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const uint32_t jj = tid * PER + i;
      loc[i] = jj < nb ? (uint32_t)__popc(s_pm[jj]) : 0u;
      sum += loc[i];
    }
    uint32_t btot;
    uint32_t e = bp_block_excl(sum, s.wtmp, &btot);
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const uint32_t jj = tid * PER + i;
      if (jj < nb) s_pre[jj] = e;
      e += loc[i];
    }
    __syncthreads();
    if (tid < n_hit) {
      const uint32_t ib = s.item_base[tid];
      if (ib >= b0 && ib < b0 + nb) s_taskcand[tid] = run + s_pre[ib - b0];
    }
    for (uint32_t jj = w; jj < nb; jj += BP_WARPS) {
      const uint32_t pm = s_pm[jj];
      if (pm & (1u << lane)) {
        const uint32_t pos = cta_base + run + s_pre[jj] + __popc(pm & ((1u << lane) - 1u));
        a.cand_pt[pos] = s_leaf[jj] * 32 + lane;
        a.cand_row[pos] = s_row[jj];
      }
    }
    run += btot;
    __syncthreads();
  }
  if (tid < n_hit && s.item_base[tid] >= n_items) s_taskcand[tid] = run;
  if (tid == 0) s_taskcand[n_hit] = run;
  __syncthreads();
  // first task of a row (level-1 node 0): the row's candidate list starts here
  const uint32_t t = blockIdx.x * a.tpc + tid;
  if (tid < a.tpc && t < a.n_tasks) {
    uint32_t row, nd;
    bool first;
    bp_task(a, t, &row, &nd, &first, bp_first_row(a));
    if (first) a.row_off[row] = cta_base + s_taskcand[rank];
  }
}

// boxes of rows [rb*n_tr, re*n_tr) must be in c->geo.box.  Asynchronous: leaves c->cand_pt / cand_row / row_off and the
// total in c->dc->n_cand (TOB_OVF_CAND set when it exceeds c->cand_cap; then nothing is written).
int broadphase(tob_ctx* c, int rb, int re, double d, int count_as) {
  return broadphase_rows(c, rb * c->n_tr, (re - rb) * c->n_tr, d, count_as);
}

// tasks per CTA: enough CTAs to cover the machine a few times over, at most BP_THREADS tasks each
static uint32_t bp_tpc(tob_ctx* c, uint32_t n_tasks) {
  uint32_t tpc = n_tasks / (4u * (uint32_t)c->sm_count);
  static int tpc_cap = -1;
  if (tpc_cap < 0) { const char* e = getenv("TRAJOPT_B200_BP_TPC"); tpc_cap = e ? atoi(e) : BP_THREADS; if (tpc_cap < 16 || tpc_cap > BP_THREADS) tpc_cap = BP_THREADS; }
  return tpc < 16u ? 16u : (tpc > (uint32_t)tpc_cap ? (uint32_t)tpc_cap : tpc);
}

void bp_args(tob_ctx* c, int row_base, int rows, double d, BpArgs& a) {
  a.box = c->geo.box.p;
  a.rows = rows; a.n1 = c->lvl[1].count; a.n_tasks = (uint32_t)rows * a.n1; a.d = d; a.row_base = (uint32_t)row_base;
  a.row_task = nullptr; a.row_l1 = nullptr;
  if (!c->cloud_n1.empty()) {          // one cloud per robot: only the level-1 nodes of the row's own cloud
    a.row_task = c->row_task.p; a.row_l1 = c->row_l1.p;
    a.n_tasks = c->h_row_task[row_base + rows] - c->h_row_task[row_base];
  }
  a.cta_row = nullptr;
  a.rows_all = (uint32_t)c->rows_all();
  for (int k = 0; k < 3; k++) {
    a.l1lo[k] = c->lvl[1].lo[k]; a.l1hi[k] = c->lvl[1].hi[k];
    a.l0lo[k] = c->lvl[0].lo[k]; a.l0hi[k] = c->lvl[0].hi[k];
  }
  a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p;
  a.bsum = c->bsum.p;
  a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p; a.row_off = c->row_off.p;
  a.cand_cap = (uint32_t)c->cand_cap;
  a.dc = c->dc.p;
  a.rec_row = c->rec_row.p; a.rec_leaf = c->rec_leaf.p; a.rec_pm = c->rec_pm.p; a.cta_ibase = c->cta_ibase.p; a.cta_nitems = c->cta_nitems.p;
  a.row_li = c->row_li.p; a.rec_cap = (uint32_t)c->rec_cap;
  // enough CTAs to cover the machine a few times over, at most BP_THREADS tasks each
  a.tpc = bp_tpc(c, a.n_tasks);
  a.cta_row = (!c->cloud_n1.empty() && row_base == 0 && rows == c->rows_all() && a.tpc == c->cta_row_tpc && c->cta_row.p) ? c->cta_row.p : nullptr;
}

// same for an arbitrary row range [row_base, row_base+rows) of geo.box (tob_box_query uses row 0 with a caller box)
int broadphase_rows(tob_ctx* c, int row_base, int rows, double d, int count_as) {
  if (c->n_pts == 0) return fail_msg(c, "broadphase: no point cloud uploaded");
  TOB_TRY(ensure_query_buffers(c));
  BpArgs a;
  bp_args(c, row_base, rows, d, a);
  const int nblk = div_up((size_t)a.n_tasks, a.tpc);
  {
    Prof prof(c, K_BP_COUNT);
    k_bp_count<<<nblk, BP_THREADS, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  {
    Prof prof(c, K_SCAN);
    k_bp_top<<<1, 1024, 0, c->stream>>>(c->bsum.p, (uint32_t)nblk, a.cand_cap, count_as, c->dc.p);
    TOB_LAUNCH_CHECK(c);
  }
  {
    Prof prof(c, K_BP_FILL);
    k_bp_fill<<<nblk, BP_THREADS, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  return 0;
}

// every buffer whose size depends only on the parameters, the cloud and the candidate capacity: allocated up front so
// that an iteration neither allocates nor reads a size back in the middle (CUDA-graph capturable)
int ensure_query_buffers(tob_ctx* c) {
  if (c->cand_cap == 0) {
    c->cand_cap = 1u << 20;
    if (const char* e = getenv("TRAJOPT_B200_CAND_CAP")) { long v = atol(e); if (v >= 256) c->cand_cap = (uint64_t)v; }   // tests: force the overflow path
  }
  const size_t rows = (size_t)c->rows_all(), U = (size_t)c->n_robots();
  const size_t n1 = c->n_levels > 1 ? c->lvl[1].count : 1;
  const size_t tasks = c->cloud_n1.empty() ? rows * n1 : (size_t)c->h_row_task[rows];
  const size_t nblk = tasks / 16 + 4 * (size_t)c->sm_count + 2;   // tasks per CTA >= 16
  const size_t self_max = c->cloud_n1.empty() ? rows * (U > 1 ? U - 1 : 0) : 0;   // independent problems have no inter-robot planes
  TOB_CUDA(c, c->bsum.ensure(nblk + 1));
  {
    // item records of the count pass: an item (hit leaf of a hit level-1 node) yields 32 x (hit fraction) candidates, so a
    // quarter of the candidate capacity is ample; a CTA that finds no room makes the fill pass repeat its tests
    const char* e = getenv("TRAJOPT_B200_BP_REC");      // 0: no records; > 1: that many (tests: most CTAs find no room)
    const long rec_on = e ? atol(e) : 1;
    c->rec_cap = rec_on ? (rec_on > 1 ? (uint64_t)rec_on : c->cand_cap / 4 + 4096) : 0;
    TOB_CUDA(c, c->rec_row.ensure(c->rec_cap + 1)); TOB_CUDA(c, c->rec_leaf.ensure(c->rec_cap + 1)); TOB_CUDA(c, c->rec_pm.ensure(c->rec_cap + 1));
    TOB_CUDA(c, c->cta_ibase.ensure(nblk + 1)); TOB_CUDA(c, c->cta_nitems.ensure(nblk + 1)); TOB_CUDA(c, c->row_li.ensure(rows + 2));
  }
  TOB_CUDA(c, c->row_off.ensure(rows + 2));
  TOB_CUDA(c, c->cand_pt.ensure(c->cand_cap + 1));
  TOB_CUDA(c, c->cand_row.ensure(c->cand_cap + 1));
  TOB_CUDA(c, c->cpl.ensure(4 * c->cand_cap + 4));
  TOB_CUDA(c, c->cflag.ensure(c->cand_cap + 1));
  TOB_CUDA(c, c->en_items.ensure((size_t)7 * rows + 8));   // barrier.cu: energy_items (EN_VMAX - 1 per row)
  TOB_CUDA(c, c->en_item_base.ensure(rows + 1));
  TOB_CUDA(c, c->gpart.ensure((size_t)54 * (7 * rows + 8)));
  TOB_CUDA(c, c->csum.ensure(c->cand_cap / 128 + 4)   /* >= chunks + 1 for any chunk size >= 128 */);
  TOB_CUDA(c, c->selfpre.ensure(rows + 2));
  TOB_CUDA(c, c->selfcnt.ensure(rows + 2));
  const size_t ob_max = c->cand_cap > c->live_cap ? c->cand_cap : c->live_cap;   // persistent mode: every live plane is listed
  TOB_CUDA(c, c->pl.ensure(4 * (ob_max + self_max) + 4));
  TOB_CUDA(c, c->pl_row.ensure(ob_max + self_max + 1));
  TOB_CUDA(c, c->pl_off.ensure(rows + 2));
  return 0;
}

}  // namespace tob
