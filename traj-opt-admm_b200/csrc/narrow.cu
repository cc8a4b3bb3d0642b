// narrow.cu -- per-pair FP64 narrowphase kernels (built with --fmad=false, see gjk.cuh):
//   * obstacle planes: 49-DOP filter -> GJK -> separating plane per (sub-segment, point) candidate
//     (Optimization3D_admm::separate_plane, Optimization3D_admm.h:69-197 with optimal_plane=0)
//   * inter-robot planes per time slot: box -> 49-DOP -> GJK -> 1-D Newton on d
//     (Optimization3D_multi::separate_self, Optimization3D_multi.h:237-342)
//   * CCD conservative step: swept 49-DOP gate + GJK on hull(P U P+sD) with the 0.8 ladder and a per-robot
//     atomicMax of the ladder exponent (Step::position_step, Step.h:21-110); inter-robot variants
//     (Step::self_step :184-256, Step::couple_self_step :112-182)
//   * packing of accepted planes into the per-row CSR the barrier kernels stream.
#include "ctx.cuh"
#include "gjk.cuh"

namespace tob {

#define TOB_MAX_LADDER TOB_LADDER

__device__ __forceinline__ void load_pts6(const double* __restrict__ src, double (*P)[3]) {
  // src: 6x3 column-major (18 doubles)
  for (int j = 0; j < 6; j++) { P[j][0] = src[j]; P[j][1] = src[j + 6]; P[j][2] = src[j + 12]; }
}

// ---- obstacle planes -------------------------------------------------------------------------------------------
struct NarrowArgs {
  uint32_t n_cand;
  const uint32_t *cand_pt, *cand_row;
  const double *px, *py, *pz;
  const double *P, *klo, *khi, *kdop;
  double dist, offset;
  double* cpl;       // n_cand x 4
  uint32_t* cflag;   // n_cand
};

__global__ void __launch_bounds__(128) k_narrow(NarrowArgs a) {
  __shared__ double s_kdop[3 * TOB_KDOP_AXES];
  for (int i = threadIdx.x; i < 3 * TOB_KDOP_AXES; i += blockDim.x) s_kdop[i] = a.kdop[i];
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  uint32_t row = a.cand_row[i], p = a.cand_pt[i];
  double q[3] = {a.px[p], a.py[p], a.pz[p]};
  uint32_t ok = 0;
  if (kdop_point_overlap(a.klo + (size_t)TOB_KDOP_AXES * row, a.khi + (size_t)TOB_KDOP_AXES * row, s_kdop, q, a.dist)) {
    double P[6][3], c[3], d;
    load_pts6(a.P + (size_t)18 * row, P);
    if (plane_point(P, q, a.dist, a.offset, c, &d)) {
      ok = 1;
      double* o = a.cpl + (size_t)4 * i;
      o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = d;
    }
  }
  a.cflag[i] = ok;
}

// ---- inter-robot planes ----------------------------------------------------------------------------------------
struct SelfArgs {
  int U, n_tr, npairs;
  const double *P, *D, *box, *klo, *khi, *kdop;
  double dist, offset, margin;
  double* self_pl;     // n_tr x npairs x 4
  uint32_t* self_ok;   // n_tr x npairs
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
  if (hit && kdop_sets_overlap(a.klo + TOB_KDOP_AXES * r0, a.khi + TOB_KDOP_AXES * r0, a.klo + TOB_KDOP_AXES * r1,
                               a.khi + TOB_KDOP_AXES * r1, a.dist)) {
    double P0[6][3], P1[6][3], c[3], d;
    load_pts6(a.P + 18 * r0, P0);
    load_pts6(a.P + 18 * r1, P1);
    if (plane_hulls(P0, P1, a.dist, c, &d)) {
      refine_d(P0, P1, c, a.offset, a.margin, &d, 10000);
      ok = 1;
      double* o = a.self_pl + (size_t)4 * t;
      o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = d;
    }
  }
  a.self_ok[t] = ok;
}

// ---- packing -----------------------------------------------------------------------------------------------------
// per row: number of accepted obstacle planes and inter-robot planes
__global__ void k_row_counts(int rows, int n_tr, int U, int npairs, int with_self, int self_begin, int self_end,
                             const uint32_t* __restrict__ row_off, const uint32_t* __restrict__ cflag_off,
                             const uint32_t* __restrict__ self_ok, uint32_t* __restrict__ row_nob,
                             uint32_t* __restrict__ row_tot) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  uint32_t nob = cflag_off[row_off[row + 1]] - cflag_off[row_off[row]];
  uint32_t ns = 0;
  if (with_self && row >= self_begin && row < self_end) {
    int u = row / n_tr, tr = row - u * n_tr;
    for (int v = 0; v < U; v++) {
      if (v == u) continue;
      int a = v < u ? v : u, b = v < u ? u : v;
      int pi = a * (U - 1) - a * (a - 1) / 2 + (b - a - 1);
      ns += self_ok[(size_t)tr * npairs + pi];
    }
  }
  row_nob[row] = nob;
  row_tot[row] = nob + ns;
}

__global__ void k_pack_obstacle(uint32_t n_cand, const uint32_t* __restrict__ cand_row, const uint32_t* __restrict__ cflag,
                                const uint32_t* __restrict__ cflag_off, const uint32_t* __restrict__ row_off,
                                const uint32_t* __restrict__ pl_off, const double* __restrict__ cpl, double* __restrict__ pl,
                                uint32_t* __restrict__ pl_row) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand || !cflag[i]) return;
  uint32_t row = cand_row[i];
  uint32_t dst = pl_off[row] + (cflag_off[i] - cflag_off[row_off[row]]);
  const double* s = cpl + (size_t)4 * i;
  double* o = pl + (size_t)4 * dst;
  o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3];
  pl_row[dst] = row;
}

// inter-robot planes go behind the obstacle planes of their row: (c, d-offset/2) for the lower robot id,
// (-c, -d-offset/2) for the higher one (Optimization3D_multi.h:300-304)
__global__ void k_pack_self(int rows, int n_tr, int U, int npairs, int self_begin, int self_end, double offset,
                            const uint32_t* __restrict__ self_ok, const double* __restrict__ self_pl,
                            const uint32_t* __restrict__ pl_off, const uint32_t* __restrict__ row_nob, double* __restrict__ pl,
                            uint32_t* __restrict__ pl_row) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows || row < self_begin || row >= self_end) return;
  int u = row / n_tr, tr = row - u * n_tr;
  uint32_t dst = pl_off[row] + row_nob[row];
  for (int v = 0; v < U; v++) {
    if (v == u) continue;
    int a = v < u ? v : u, b = v < u ? u : v;
    int pi = a * (U - 1) - a * (a - 1) / 2 + (b - a - 1);
    size_t t = (size_t)tr * npairs + pi;
    if (!self_ok[t]) continue;
    const double* s = self_pl + 4 * t;
    double* o = pl + (size_t)4 * dst;
    if (u == a) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3] - 0.5 * offset; }
    else { o[0] = -s[0]; o[1] = -s[1]; o[2] = -s[2]; o[3] = -s[3] - 0.5 * offset; }
    pl_row[dst] = row;
    dst++;
  }
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
  a.self_pl = c->self_pl.p; a.self_ok = c->self_ok.p;
  if (n) {
    Prof prof(c, K_SELF_PLANES);
    k_self_planes<<<div_up(n, 64), 64, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  c->ctr.self_pairs += n;
  return 0;
}

// shared tail: per-row totals -> CSR offsets -> scatter.  have_cand: obstacle planes come from the candidate scratch
static int pack_rows(tob_ctx* c, int rb, int re, bool have_cand, bool ws) {
  cudaStream_t st = c->stream;
  const int rows = c->rows_all(), U = c->n_robots(), npairs = U * (U - 1) / 2;
  uint64_t nc = have_cand ? c->n_cand : 0;
  TOB_CUDA(c, c->row_nob.ensure(rows + 1));
  TOB_CUDA(c, c->row_ntot.ensure(rows + 1));
  TOB_CUDA(c, c->pl_off.ensure(rows + 2));
  TOB_CUDA(c, c->self_ok.ensure(1));
  k_row_counts<<<div_up(rows, 128), 128, 0, st>>>(rows, c->n_tr, U, npairs, ws ? 1 : 0, rb * c->n_tr, re * c->n_tr,
                                                  c->row_off.p, c->cflag_off.p, c->self_ok.p, c->row_nob.p, c->row_ntot.p);
  TOB_LAUNCH_CHECK(c);
  uint32_t* tot_dev = (uint32_t*)c->red.p;
  TOB_TRY(exclusive_scan_u32(c, c->row_ntot.p, c->pl_off.p, rows, tot_dev));
  uint32_t* hp = (uint32_t*)c->h_pinned;
  TOB_CUDA(c, cudaMemcpyAsync(hp, tot_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  uint64_t np = hp[0];
  c->n_planes = np;
  TOB_CUDA(c, c->pl.ensure(4 * np + 4));
  TOB_CUDA(c, c->pl_row.ensure(np + 1));
  if (nc) {
    Prof prof(c, K_PACK);
    k_pack_obstacle<<<div_up(nc, 256), 256, 0, st>>>((uint32_t)nc, c->cand_row.p, c->cflag.p, c->cflag_off.p, c->row_off.p,
                                                      c->pl_off.p, c->cpl.p, c->pl.p, c->pl_row.p);
    TOB_LAUNCH_CHECK(c);
  }
  if (ws) {
    k_pack_self<<<div_up(rows, 128), 128, 0, st>>>(rows, c->n_tr, U, npairs, rb * c->n_tr, re * c->n_tr, c->prm.offset,
                                                   c->self_ok.p, c->self_pl.p, c->pl_off.p, c->row_nob.p, c->pl.p, c->pl_row.p);
    TOB_LAUNCH_CHECK(c);
  }
  c->ctr.planes += np;
  return 0;
}

// candidates (c->cand_*, c->row_off, c->n_cand) + row geometry must be current.  Leaves the packed plane CSR.
// Inter-robot planes need geo.P/box/klo/khi of ALL robots.
int narrowphase_planes(tob_ctx* c, int rb, int re, int with_self) {
  cudaStream_t st = c->stream;
  uint64_t nc = c->n_cand;
  TOB_CUDA(c, c->cpl.ensure(4 * nc + 4));
  TOB_CUDA(c, c->cflag.ensure(nc + 1));
  TOB_CUDA(c, c->cflag_off.ensure(nc + 2));
  if (nc) {
    NarrowArgs a;
    a.n_cand = (uint32_t)nc; a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p;
    a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p;
    a.P = c->geo.P.p; a.klo = c->geo.klo.p; a.khi = c->geo.khi.p; a.kdop = c->d_kdop.p;
    a.dist = c->prm.offset + c->prm.margin; a.offset = c->prm.offset;
    a.cpl = c->cpl.p; a.cflag = c->cflag.p;
    Prof prof(c, K_NARROW);
    k_narrow<<<div_up(nc, 128), 128, 0, st>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  TOB_TRY(exclusive_scan_u32(c, c->cflag.p, c->cflag_off.p, nc, nullptr));
  bool ws = with_self && c->n_robots() > 1;
  if (ws) TOB_TRY(self_planes(c));
  c->ctr.dcd_candidates += nc;
  return pack_rows(c, rb, re, true, ws);
}

// inter-robot planes only (Optimization3D_multi::separate_self on empty lists): geo of ALL robots must be current
int pack_self_only(tob_ctx* c) {
  const int rows = c->rows_all();
  TOB_CUDA(c, c->row_off.ensure(rows + 2));
  TOB_CUDA(c, c->cflag_off.ensure(2));
  TOB_CUDA(c, cudaMemsetAsync(c->row_off.p, 0, (rows + 2) * sizeof(uint32_t), c->stream));
  TOB_CUDA(c, cudaMemsetAsync(c->cflag_off.p, 0, 2 * sizeof(uint32_t), c->stream));
  c->n_cand = 0;
  TOB_TRY(self_planes(c));
  return pack_rows(c, 0, c->n_robots(), false, true);
}

// caller-provided plane lists (the reference's c_lists/d_lists) for robots [rb,re): offsets over (re-rb)*n_tr rows
int pack_planes_from_host(tob_ctx* c, int rb, int re, const uint32_t* offsets, const double* cc, const double* dd) {
  cudaStream_t st = c->stream;
  const int rows = c->rows_all(), nloc = (re - rb) * c->n_tr;
  uint64_t np = offsets[nloc];
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
  TOB_CUDA(c, c->pl_off.ensure(rows + 2));
  TOB_CUDA(c, c->pl.ensure(4 * np + 4));
  TOB_CUDA(c, c->pl_row.ensure(np + 1));
  TOB_CUDA(c, cudaMemcpyAsync(c->pl_off.p, off.data(), (rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(c->pl.p, pl.data(), (4 * np + 4) * sizeof(double), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaMemcpyAsync(c->pl_row.p, prow.data(), (np + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TOB_CUDA(c, cudaStreamSynchronize(st));
  c->n_planes = np;
  return 0;
}

// ---- CCD ladder vs the cloud -----------------------------------------------------------------------------------
struct CcdArgs {
  uint32_t n_cand;
  const uint32_t *cand_pt, *cand_row;
  const double *px, *py, *pz;
  const double *P, *D, *klo, *khi, *kdop, *steps;
  double offset;
  int n_tr, first_robot;
  int* kmax;   // per robot
};

__device__ __forceinline__ void moved_points(const double (*P)[3], const double (*D)[3], double s, double (*A)[3]) {
  for (int j = 0; j < 6; j++)
    for (int k = 0; k < 3; k++) {
      A[j][k] = P[j][k] + 0.0 * D[j][k];
      A[j + 6][k] = P[j][k] + s * D[j][k];
    }
}

__global__ void __launch_bounds__(128) k_ccd(CcdArgs a) {
  __shared__ double s_kdop[3 * TOB_KDOP_AXES];
  for (int i = threadIdx.x; i < 3 * TOB_KDOP_AXES; i += blockDim.x) s_kdop[i] = a.kdop[i];
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  uint32_t row = a.cand_row[i], p = a.cand_pt[i];
  int robot = row / a.n_tr;
  double q[1][3] = {{a.px[p], a.py[p], a.pz[p]}};
  double P[6][3], D[6][3], A[12][3];
  load_pts6(a.P + (size_t)18 * row, P);
  load_pts6(a.D + (size_t)18 * row, D);
  int k = *((volatile int*)(a.kmax + robot));
  double s = a.steps[k];
  moved_points(P, D, s, A);
  // CCD::KDOPCCD(P, D, q, offset, 0, step)
  if (!kdop_overlap<12, 1>(A, q, s_kdop, a.offset)) return;
  const double d2 = a.offset * a.offset;
  while (k < TOB_MAX_LADDER) {
    double v[3];
    gjk_witness<12, 1>(A, q, v);
    double dist2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (!(dist2 <= d2)) break;
    k++;
    s = a.steps[k];
    moved_points(P, D, s, A);
  }
  atomicMax(a.kmax + robot, k);
}

// c->cand_* must hold the swept-box candidates (global rows); c->kmax (per robot) must be zeroed by the caller.
int ccd_position_steps(tob_ctx* c) {
  uint64_t nc = c->n_cand;
  if (nc) {
    CcdArgs a;
    a.n_cand = (uint32_t)nc; a.cand_pt = c->cand_pt.p; a.cand_row = c->cand_row.p;
    a.px = c->px.p; a.py = c->py.p; a.pz = c->pz.p;
    a.P = c->geo.P.p; a.D = c->geo.D.p; a.klo = c->geo.klo.p; a.khi = c->geo.khi.p; a.kdop = c->d_kdop.p;
    a.steps = c->d_steps.p; a.offset = c->prm.offset; a.n_tr = c->n_tr; a.first_robot = 0;
    a.kmax = c->kmax.p;
    Prof prof(c, K_CCD);
    k_ccd<<<div_up(nc, 128), 128, 0, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  c->ctr.ccd_candidates += nc;
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

// phase 2 (one thread, sequential like the reference): resolve the colliding pairs in (slot, pair) order.
// The list is short (pairs that really collide when both robots take their full Newton step); it is sorted first so
// the result does not depend on the order the filter threads appended it.
__global__ void k_self_ccd_resolve(SelfCcdArgs a, double* steps_out, int* overflow) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int u = 0; u < a.U; u++) steps_out[u] = 1.0;
  int kshared = 0;
  uint32_t nh = a.hit[a.cap];
  if (nh > a.cap) { *overflow = 1; nh = a.cap; }
  for (uint32_t i = 1; i < nh; i++) {          // insertion sort
    uint32_t v = a.hit[i];
    int j = (int)i - 1;
    while (j >= 0 && a.hit[j] > v) { a.hit[j + 1] = a.hit[j]; j--; }
    a.hit[j + 1] = v;
  }
  for (uint32_t i = 0; i < nh; i++) {
    {
      const int tr = a.hit[i] / a.npairs, pi = a.hit[i] - tr * a.npairs;
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
      while (guard++ < TOB_MAX_LADDER) {
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
  }
  if (a.coupled) steps_out[0] = a.steps[kshared];
}

// geo.P / geo.D of ALL robots must be current (compute_rows with mode 2).  steps_dev: n_robots (or [0] coupled).
int self_ccd_steps(tob_ctx* c, int coupled, double* steps_dev) {
  const int U = c->n_robots();
  int npairs = U * (U - 1) / 2;
  size_t n = (size_t)c->n_tr * npairs;
  const uint32_t cap = 16384;
  TOB_CUDA(c, c->self_hits.ensure(cap + 2));
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
  k_self_ccd_resolve<<<1, 32, 0, c->stream>>>(a, steps_dev, (int*)(c->self_hits.p + cap + 1));
  TOB_LAUNCH_CHECK(c);
  c->ctr.self_pairs += n;
  return 0;
}

}  // namespace tob
