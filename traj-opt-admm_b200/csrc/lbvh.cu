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
// Broadphase work decomposition: one task = (row, level-1 node).  A warp owns 32 consecutive tasks: every lane
// pre-tests its task's node box, then the warp walks the hit tasks cooperatively (32 lanes = 32 leaf boxes, then
// 32 lanes = 32 points of a hit leaf).  Two passes (count, exclusive scan, fill) give a deterministic candidate
// order (row, Morton position) without atomics.
//
// Algorithmic bytes: build 128 B/point; query 48 B/row + 16 B/(row x L1 node) + 28 B/candidate.
#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"

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

__global__ void k_morton(const double* __restrict__ V, uint32_t n, double lx, double ly, double lz, double sx, double sy,
                         double sz, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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

int lbvh_build(tob_ctx* c, const double* V_host, uint32_t n) {
  if (n == 0) return fail_msg(c, "tob_cloud_upload: empty cloud");
  cudaStream_t st = c->stream;
  uint32_t n_pad = (n + 31u) & ~31u;
  DBuf<double> V;
  DBuf<uint64_t> key, key2;
  DBuf<uint32_t> idx, idx2;
  DBuf<double> part;
  DBuf<uint8_t> tmp;
  TOB_CUDA(c, V.ensure((size_t)3 * n));
  TOB_CUDA(c, cudaMemcpyAsync(V.p, V_host, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  const int nb = 296;
  TOB_CUDA(c, part.ensure(nb * 6));
  k_minmax<<<nb, 256, 0, st>>>(V.p, n, part.p);
  TOB_LAUNCH_CHECK(c);
  std::vector<double> hp(nb * 6);
  TOB_CUDA(c, cudaMemcpyAsync(hp.data(), part.p, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int b = 0; b < nb; b++)
    for (int a = 0; a < 3; a++) {
      lo[a] = fmin(lo[a], hp[b * 6 + a]);
      hi[a] = fmax(hi[a], hp[b * 6 + 3 + a]);
    }
  double sc[3];
  for (int a = 0; a < 3; a++) {
    double ext = hi[a] - lo[a];
    sc[a] = ext > 0 ? 2097151.0 / ext : 0.0;
  }
  TOB_CUDA(c, key.ensure(n)); TOB_CUDA(c, key2.ensure(n));
  TOB_CUDA(c, idx.ensure(n)); TOB_CUDA(c, idx2.ensure(n));
  k_morton<<<div_up(n, 256), 256, 0, st>>>(V.p, n, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2], key.p, idx.p);
  TOB_LAUNCH_CHECK(c);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key2.p, idx.p, idx2.p, (int)n, 0, 63, st);
  TOB_CUDA(c, tmp.ensure(tmp_bytes));
  TOB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key.p, key2.p, idx.p, idx2.p, (int)n, 0, 63, st));

  TOB_CUDA(c, c->px.ensure(n_pad)); TOB_CUDA(c, c->py.ensure(n_pad)); TOB_CUDA(c, c->pz.ensure(n_pad));
  TOB_CUDA(c, c->pid.ensure(n_pad));
  k_gather<<<div_up(n_pad, 256), 256, 0, st>>>(V.p, n, n_pad, idx2.p, c->px.p, c->py.p, c->pz.p, c->pid.p);
  TOB_LAUNCH_CHECK(c);

  // levels: 0 = leaves (32 points), then 32-ary up to a single node
  uint32_t counts[TOB_MAX_LEVELS], pads[TOB_MAX_LEVELS];
  int nl = 0;
  uint32_t cnt = n_pad / 32;
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
    uint32_t n_child = (l == 0) ? n_pad : counts[l - 1];
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
  c->n_pts = n; c->n_pad = n_pad;
  V.release(); key.release(); key2.release(); idx.release(); idx2.release(); part.release(); tmp.release();
  return 0;
}

// ---- exclusive scan (uint32) ---------------------------------------------------------------------------------
#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sm /*>=32*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
    uint32_t si = s;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    sm[lane] = si - s;
    if (lane == 31) *total = si;
  }
  __syncthreads();
  uint32_t r = sm[w] + inc - v;
  __syncthreads();
  return r;
}

__global__ void k_scan_reduce(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ bsum) {
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t k = base + (size_t)i * SCAN_BLOCK + threadIdx.x;
    if (k < n) s += in[k];
  }
  __shared__ uint32_t sm[32];
  __shared__ uint32_t tot;
  block_excl_scan(s, &tot, sm);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: exclusive scan of bsum[nb] in place, total to bsum[nb]
__global__ void k_scan_top(uint32_t* bsum, uint32_t nb) {
  __shared__ uint32_t sm[32];
  __shared__ uint32_t tot;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += blockDim.x) {
    uint32_t k = base + threadIdx.x;
    uint32_t v = k < nb ? bsum[k] : 0;
    uint32_t e = block_excl_scan(v, &tot, sm);
    if (k < nb) bsum[k] = carry + e;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

__global__ void k_scan_down(const uint32_t* __restrict__ in, size_t n, const uint32_t* __restrict__ bsum, uint32_t nb,
                            uint32_t* __restrict__ out, uint32_t* __restrict__ total_dev) {
  __shared__ uint32_t sm[32];
  __shared__ uint32_t tot;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  uint32_t e = block_excl_scan(s, &tot, sm) + bsum[blockIdx.x];
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < n) out[base + i] = e;
    e += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[n] = bsum[nb];
    if (total_dev) *total_dev = bsum[nb];
  }
}

int exclusive_scan_u32(tob_ctx* c, const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_dev) {
  uint32_t nb = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
  if (nb == 0) nb = 1;
  TOB_CUDA(c, c->scan_tmp.ensure(nb + 1));
  Prof prof(c, K_SCAN);   // the three scan kernels are timed as one unit
  k_scan_reduce<<<nb, SCAN_BLOCK, 0, c->stream>>>(in, n, c->scan_tmp.p);
  TOB_LAUNCH_CHECK(c);
  k_scan_top<<<1, 1024, 0, c->stream>>>(c->scan_tmp.p, nb);
  TOB_LAUNCH_CHECK(c);
  k_scan_down<<<nb, SCAN_BLOCK, 0, c->stream>>>(in, n, c->scan_tmp.p, nb, out, total_dev);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// ---- broadphase ------------------------------------------------------------------------------------------------
struct BpArgs {
  const double* box;           // rows x 6
  uint32_t rows, n1, n_tasks, row_base;
  double d;
  const double *l1lo[3], *l1hi[3], *l0lo[3], *l0hi[3];
  const double *px, *py, *pz;
  uint32_t* task_cnt;          // count pass: out
  const uint32_t* task_off;    // fill pass: in
  uint32_t *cand_pt, *cand_row;
};

// the reference predicate with the query box [qlo,qhi] as "this" and the node/point as the argument
__device__ __forceinline__ bool box_hit(double nlo, double nhi, double qlo, double qhi, double d) {
  return !(nhi + d < qlo) && !(nlo > qhi + d);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_broadphase(BpArgs a) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t t = warp * 32 + lane;
  bool hit = false;
  if (t < a.n_tasks) {
    uint32_t r = t / a.n1, nd = t - r * a.n1;
    const double* q = a.box + (size_t)6 * (a.row_base + r);
    hit = box_hit(a.l1lo[0][nd], a.l1hi[0][nd], q[0], q[3], a.d) && box_hit(a.l1lo[1][nd], a.l1hi[1][nd], q[1], q[4], a.d) &&
          box_hit(a.l1lo[2][nd], a.l1hi[2][nd], q[2], q[5], a.d);
    if (!FILL && !hit) a.task_cnt[t] = 0;
  }
  uint32_t mask = __ballot_sync(0xffffffffu, hit);
  while (mask) {
    uint32_t b = __ffs(mask) - 1;
    mask &= mask - 1;
    uint32_t tt = warp * 32 + b;
    uint32_t r = tt / a.n1, nd = tt - r * a.n1;
    const double* q = a.box + (size_t)6 * (a.row_base + r);
    double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5];
    uint32_t leaf = nd * 32 + lane;   // level-0 arrays are padded to 32 with empty boxes
    bool lh = box_hit(a.l0lo[0][leaf], a.l0hi[0][leaf], q0, q3, a.d) && box_hit(a.l0lo[1][leaf], a.l0hi[1][leaf], q1, q4, a.d) &&
              box_hit(a.l0lo[2][leaf], a.l0hi[2][leaf], q2, q5, a.d);
    uint32_t lmask = __ballot_sync(0xffffffffu, lh);
    uint32_t cnt = 0;
    uint32_t base = FILL ? a.task_off[tt] : 0;
    while (lmask) {
      uint32_t lb = __ffs(lmask) - 1;
      lmask &= lmask - 1;
      uint32_t p = (nd * 32 + lb) * 32 + lane;
      double x = a.px[p], y = a.py[p], z = a.pz[p];
      bool ok = box_hit(x, x, q0, q3, a.d) && box_hit(y, y, q1, q4, a.d) && box_hit(z, z, q2, q5, a.d);
      uint32_t pm = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) {
        uint32_t pos = base + cnt + __popc(pm & ((1u << lane) - 1u));
        a.cand_pt[pos] = p;
        a.cand_row[pos] = a.row_base + r;
      }
      cnt += __popc(pm);
    }
    if (!FILL && lane == 0) a.task_cnt[tt] = cnt;
  }
}

// candidate offsets for ALL rows of the context; rows outside the queried range [row_base, row_base+rows) are empty
__global__ void k_row_offsets(const uint32_t* __restrict__ task_off, uint32_t rows, uint32_t row_base, uint32_t rows_all,
                              uint32_t n1, uint32_t* __restrict__ row_off) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > rows_all) return;
  uint32_t lr = g < row_base ? 0u : (g - row_base > rows ? rows : g - row_base);
  row_off[g] = task_off[(size_t)lr * n1];
}

// boxes of rows [rb*n_tr, re*n_tr) must be in c->geo.box.  Leaves c->cand_pt / cand_row / row_off, total on host.
int broadphase(tob_ctx* c, int rb, int re, double d, uint64_t* total_host) {
  return broadphase_rows(c, rb * c->n_tr, (re - rb) * c->n_tr, d, total_host);
}

// same for an arbitrary row range [row_base, row_base+rows) of geo.box (tob_box_query uses row 0 with a caller box)
int broadphase_rows(tob_ctx* c, int row_base, int rows, double d, uint64_t* total_host) {
  if (c->n_pts == 0) return fail_msg(c, "broadphase: no point cloud uploaded");
  BpArgs a;
  a.box = c->geo.box.p;
  a.rows = rows; a.n1 = c->lvl[1].count; a.n_tasks = (uint32_t)rows * a.n1; a.d = d; a.row_base = (uint32_t)row_base;
  for (int k = 0; k < 3; k++) {
    a.l1lo[k] = c->lvl[1].lo[k]; a.l1hi[k] = c->lvl[1].hi[k];
    a.l0lo[k] = c->lvl[0].lo[k]; a.l0hi[k] = c->lvl[0].hi[k];
  }
  a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p;
  TOB_CUDA(c, c->task_cnt.ensure(a.n_tasks + 1));
  TOB_CUDA(c, c->task_off.ensure(a.n_tasks + 1));
  TOB_CUDA(c, c->row_off.ensure(c->rows_all() + 2));
  a.task_cnt = c->task_cnt.p; a.task_off = c->task_off.p;
  a.cand_pt = nullptr; a.cand_row = nullptr;
  int nblk = div_up((size_t)a.n_tasks, 256);
  {
    Prof prof(c, K_BP_COUNT);
    k_broadphase<false><<<nblk, 256, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  uint32_t* tot_dev = (uint32_t*)c->red.p;
  TOB_TRY(exclusive_scan_u32(c, c->task_cnt.p, c->task_off.p, a.n_tasks, tot_dev));
  uint32_t* hp = (uint32_t*)c->h_pinned;
  TOB_CUDA(c, cudaMemcpyAsync(hp, tot_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TOB_CUDA(c, cudaStreamSynchronize(c->stream));
  uint64_t total = hp[0];
  c->n_cand = total;
  *total_host = total;
  TOB_CUDA(c, c->cand_pt.ensure(total + 1));
  TOB_CUDA(c, c->cand_row.ensure(total + 1));
  a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p;
  if (total) {
    Prof prof(c, K_BP_FILL);
    k_broadphase<true><<<nblk, 256, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  k_row_offsets<<<div_up(c->rows_all() + 1, 256), 256, 0, c->stream>>>(c->task_off.p, rows, a.row_base, c->rows_all(), a.n1,
                                                                       c->row_off.p);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

}  // namespace tob
