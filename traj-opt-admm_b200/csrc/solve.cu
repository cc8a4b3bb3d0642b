// solve.cu -- ADMM primal Newton solve and slack / dual update.
//
// Primal (replaces Optimization3D_admm::spline_descent_direction, Optimization3D_admm.h:400-503, and the
// multi-robot variant Optimization3D_multi.h:659-752): the per-piece 19x19 blocks overlap by 9 unknowns
// (stride 3 control points), so after dropping the 4 fixed end control points the Newton matrix is banded with
// half-bandwidth 17 plus one dense arrow row/column (the shared piece time).  The reference builds the dense
// (3T+1)^2 matrix, calls sparseView() and Eigen::SimplicialLLT; here one CTA per robot assembles the band in
// shared memory, factors it with a right-looking banded Cholesky and eliminates the arrow by a Schur complement:
//     [B a; a^T h][x; t] = -[g; gt],  y = B^-1 a, z = B^-1 g,  t = (a.z - gt)/(h - a.y),  x = -z - y t.
// Same solution as the reference up to round-off (different elimination order).
//
// Slack / dual (replaces Optimization3D_admm::update_slack_lambda :231-398, Gradient_admm::slack_gradient
// :574-622, dynamic_gradient :633-671, Energy_admm::slack_energy :172-190): one thread per (robot, piece),
// 19x19 (13x13 at the two ends) dense Newton step + 0.8 backtracking + dual ascent.
#include "ctx.cuh"
#include "dense.cuh"

namespace tob {

#define BW 18   // band storage width: diagonal + 17 sub-diagonals

struct SolveArgs {
  const double *pc_g, *pc_h;
  int P, T, robot_begin, dense_shift;
  double* dir;       // robots x 3T
  double* tdir;      // robots
  double* wolfe;     // robots
  double* gnorm;     // robots
  int* status;       // robots: 0 ok, 1 not SPD
  double* gband;     // optional global workspace (robots x m x 22) when shared memory is too small
  int use_global;
  DevCounts* dc;     // a failed factorisation sets TOB_ERR_SOLVE: the iteration is not committed
};

__global__ void __launch_bounds__(256) k_solve(SolveArgs a) {
  extern __shared__ double sm[];
  const int robot = a.robot_begin + blockIdx.x;
  const int m = 3 * (a.T - 4);           // banded unknowns
  double* Bd = a.use_global ? a.gband + (size_t)blockIdx.x * m * 22 : sm;   // m x BW  : Bd[r*BW + k] = A(r, r-k)
  double* va = Bd + (size_t)m * BW;      // arrow column -> y
  double* vg = va + m;                   // gradient     -> z
  double* g0 = vg + m;                   // copy of the gradient
  double* gx = g0 + m;                   // unused spare
  __shared__ double s_h, s_gt;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  for (int i = tid; i < m * BW; i += blockDim.x) Bd[i] = 0;
  for (int i = tid; i < m; i += blockDim.x) { va[i] = 0; vg[i] = 0; }
  if (tid == 0) { s_h = 0; s_gt = 0; s_fail = 0; }
  __syncthreads();
  // assembly: piece sp covers full coordinates 9*sp .. 9*sp+17, reduced index = full - 6.
  // Every (r,c) entry receives at most two piece contributions: add them in piece order (deterministic).
  for (int r = tid; r < m; r += blockDim.x) {
    int f = r + 6;
    int sp_hi = f / 9; if (sp_hi > a.P - 1) sp_hi = a.P - 1;
    int sp_lo = (f - 17 + 8) / 9; if (f - 17 < 0) sp_lo = 0;
    double ga = 0, aa = 0;
    for (int sp = sp_lo; sp <= sp_hi; sp++) {
      int lr = f - 9 * sp;
      if (lr < 0 || lr > 17) continue;
      const double* H = a.pc_h + (size_t)361 * ((size_t)robot * a.P + sp);
      ga += a.pc_g[(size_t)19 * ((size_t)robot * a.P + sp) + lr];
      aa += H[lr + 19 * 18];
      for (int k = 0; k < BW; k++) {
        int cidx = r - k;
        if (cidx < 0) break;
        int lc = cidx + 6 - 9 * sp;
        if (lc < 0 || lc > 17) continue;
        Bd[r * BW + k] += H[lr + 19 * lc];
      }
    }
    vg[r] = ga; g0[r] = ga; va[r] = aa;
  }
  if (tid == 0) {
    double h = 0, gt = 0;
    for (int sp = 0; sp < a.P; sp++) {
      h += a.pc_h[(size_t)361 * ((size_t)robot * a.P + sp) + 18 + 19 * 18];
      gt += a.pc_g[(size_t)19 * ((size_t)robot * a.P + sp) + 18];
    }
    s_h = h; s_gt = gt;
  }
  __syncthreads();
  // right-looking banded Cholesky by ONE warp (no block barriers on the 3(T-4)-long dependency chain), in place:
  // Bd becomes L (L(r, r-k) at Bd[r*BW+k]).  The two right-hand sides (arrow column va, gradient vg) ride along as
  // extra rows, so the forward substitution L w = b is finished when the factorisation is.
  const int lane = tid & 31, wp = tid >> 5;
  if (wp == 0) {
    int fail = 0;
    // the diagonal stores 1/L(j,j): every division on the 3(T-4)-long dependency chain becomes a multiply
    double piv = Bd[0];
    if (!(piv > 0)) { fail = 1; piv = 1; }
    double rinv = rsqrt(piv);
    rinv = rinv * (1.5 - 0.5 * piv * rinv * rinv);            // one Newton step: full double accuracy
    for (int j = 0; j < m; j++) {
      const int cnt = (m - 1 - j) < 17 ? (m - 1 - j) : 17;   // rows below the pivot inside the band
      double l = 0;
      if (lane < cnt) l = Bd[(j + 1 + lane) * BW + lane + 1] * rinv;
      const double ya = va[j] * rinv, yg = vg[j] * rinv;
      // look-ahead: the next pivot only needs l of lane 0, start its reciprocal square root before the trailing update
      const double l0 = __shfl_sync(0xffffffffu, l, 0);
      double rinv_next = 0;
      if (j + 1 < m) {
        double pn = Bd[(j + 1) * BW] - l0 * l0;
        if (!(pn > 0)) { fail = 1; pn = 1; }
        rinv_next = rsqrt(pn);
        rinv_next = rinv_next * (1.5 - 0.5 * pn * rinv_next * rinv_next);
      }
      __syncwarp();
      if (lane < cnt) Bd[(j + 1 + lane) * BW + lane + 1] = l;
      if (lane == 0) { Bd[j * BW] = rinv; va[j] = ya; vg[j] = yg; }
      // trailing update A(j+1+p, j+1+q) -= L(j+1+p, j) L(j+1+q, j), q <= p: lane p owns row j+1+p
      // fully unrolled so the 17 shared-memory read-modify-writes of a lane are independent instructions
      {
        double* rowp = Bd + (j + 1 + lane) * BW + lane;
#pragma unroll
        for (int q = 0; q < 17; q++) {
          double lq = __shfl_sync(0xffffffffu, l, q);
          if (lane < cnt && q <= lane) rowp[-q] -= l * lq;
        }
      }
      if (lane < cnt) { va[j + 1 + lane] -= l * ya; vg[j + 1 + lane] -= l * yg; }
      rinv = rinv_next;
      __syncwarp();
    }
    if (lane == 0 && fail) s_fail = 1;
    // backward substitution L^T x = w for both right-hand sides: lane k-1 holds the k-th sub-diagonal term
    for (int r = m - 1; r >= 0; r--) {
      double sa = 0, sg = 0;
      int k = lane + 1;
      if (k < BW && r + k < m) { double lv = Bd[(r + k) * BW + k]; sa = lv * va[r + k]; sg = lv * vg[r + k]; }
      for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sg += __shfl_xor_sync(0xffffffffu, sg, o); }
      if (lane == 0) { double dd = Bd[r * BW]; va[r] = (va[r] - sa) * dd; vg[r] = (vg[r] - sg) * dd; }
      __syncwarp();
    }
  }
  __syncthreads();
  // Schur complement for the arrow, then direction / wolfe / gnorm
  __shared__ double s_red[8][4];
  double ay = 0, az = 0, gg = 0;
  (void)gx;
  // a . y and a . z need the ORIGINAL arrow column: recompute it from the piece blocks (cheap)
  for (int r = tid; r < m; r += blockDim.x) {
    int f = r + 6;
    double aa = 0;
    for (int sp = 0; sp < a.P; sp++) {
      int lr = f - 9 * sp;
      if (lr < 0 || lr > 17) continue;
      aa += a.pc_h[(size_t)361 * ((size_t)robot * a.P + sp) + lr + 19 * 18];
    }
    ay += aa * va[r];
    az += aa * vg[r];
    gg += g0[r] * g0[r];
  }
  for (int o = 16; o; o >>= 1) {
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    az += __shfl_xor_sync(0xffffffffu, az, o);
    gg += __shfl_xor_sync(0xffffffffu, gg, o);
  }
  if (lane == 0) { s_red[wp][0] = ay; s_red[wp][1] = az; s_red[wp][2] = gg; }
  __syncthreads();
  __shared__ double s_t;
  if (tid == 0) {
    double AY = 0, AZ = 0, GG = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { AY += s_red[i][0]; AZ += s_red[i][1]; GG += s_red[i][2]; }
    double schur = s_h - AY;
    if (!(schur > 0)) s_fail = 1;
    double t = (AZ - s_gt) / schur;
    s_t = t;
    a.tdir[robot] = t;
    a.gnorm[robot] = sqrt(GG + s_gt * s_gt);
  }
  __syncthreads();
  const double t = s_t;
  double wl = 0;
  double* dir = a.dir + (size_t)robot * 3 * a.T;
  for (int r = tid; r < m; r += blockDim.x) {
    double x = -vg[r] - va[r] * t;
    wl += x * g0[r];
    int mm = r / 3, k = r - 3 * mm;
    dir[(size_t)k * a.T + 2 + mm] = x;
  }
  for (int i = tid; i < 12; i += blockDim.x) {   // fixed end control points: rows 0,1,T-2,T-1
    int k = i / 4, w = i % 4;
    int rr = w < 2 ? w : a.T - 4 + w;
    dir[(size_t)k * a.T + rr] = 0.0;
  }
  for (int o = 16; o; o >>= 1) wl += __shfl_xor_sync(0xffffffffu, wl, o);
  __syncthreads();
  if (lane == 0) s_red[wp][3] = wl;
  __syncthreads();
  if (tid == 0) {
    double W = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) W += s_red[i][3];
    W += t * s_gt;
    a.wolfe[robot] = -W;
    a.status[robot] = s_fail;
    if (s_fail && a.dc && !(a.dc->overflow & TOB_OVF_RETRY)) atomicOr(&a.dc->overflow, TOB_ERR_SOLVE);
  }
}

// ---- block cyclic reduction ------------------------------------------------------------------------------------------
// The banded matrix is block tridiagonal with N = P+1 blocks of 9 unknowns (3 control points): piece sp couples blocks
// sp and sp+1.  Sequential Cholesky has a dependency chain of 3(T-4) pivots (~1 M cycles on one warp at P=64); cyclic
// reduction eliminates every other block per level, all blocks of a level in parallel (one warp per block):
//   eliminated block i (neighbours i-s, i+s):  D_i = L L^T,  U = L^-1 A(i,i-s),  V = L^-1 A(i,i+s),  c = L^-1 r_i
//   surviving block j:  D_j -= V^T V (from j-s) + U^T U (from j+s),  r_j -= V^T c + U^T c,
//                       A(j,j-2s) = -V^T U (from j-s),  A(j,j+2s) = -U^T V (from j+s)
// log2(N) levels, then x_i = L^-T (c - U x_{i-s} - V x_{i+s}) in reverse.  SPD Schur complements stay SPD.
// The 4 fixed end control points are kept as identity rows (x = 0) so every block has size 9.
#define BS 9
#define BLK (3 * 81 + 18 + 9)     // El, D, Eu (81 each), R (9x2), 1/diag(L) (9)

// 1/sqrt(x).  CUDA's double rsqrt() is within 1-2 ulp; a further Newton step would add three dependent FP64 operations
// (~120 cycles) to every pivot of the block Cholesky, whose dependency chain IS the latency of the solve.
__device__ __forceinline__ double rsqrt_full(double x) { return rsqrt(x); }

// in-place lower Cholesky of the 9x9 col-major matrix D by one warp; dinv receives 1/L(k,k).  Returns false on a
// non-positive pivot (same value on every lane).  Lane j keeps column j in registers (static indices, fully unrolled); per
// pivot the owner lane scales its column into shared memory, one warp barrier, and the lanes to its right update their
// columns from it: one barrier and one shared-memory round trip per pivot, no index arithmetic.
__device__ inline bool warp_chol9(double* D, double* dinv) {
  const int lane = threadIdx.x & 31;
  double c[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) c[i] = lane < BS ? D[i + BS * lane] : 0.0;
  int ok = 1;
#pragma unroll
  for (int k = 0; k < BS; k++) {
    if (lane == k) {
      double p = c[k];
      if (!(p > 0)) { ok = 0; p = 1; }
      const double ri = rsqrt_full(p);
      D[k + BS * k] = p * ri;
      dinv[k] = ri;
#pragma unroll
      for (int i = k + 1; i < BS; i++) D[i + BS * k] = c[i] * ri;
    }
    __syncwarp();
    if (lane > k && lane < BS) {
      const double ljk = D[lane + BS * k];
#pragma unroll
      for (int i = k + 1; i < BS; i++)
        if (i >= lane) c[i] -= D[i + BS * k] * ljk;
    }
  }
  __syncwarp();
  return __all_sync(0xffffffffu, ok) != 0;
}

struct BcrArgs {
  const double *pc_g, *pc_h;
  int P, T, robot_begin;
  double *dir, *tdir, *wolfe, *gnorm;
  int* status;
  // coupled multi-robot system (Optimization3D_multi::update_spline :508-639): block diagonal per robot + ONE shared
  // arrow row for the common piece time.  Each robot's CTA exports z = B^-1 g, y = B^-1 a and its partial sums; the
  // Schur complement over all robots is closed by k_couple_finish.  Null for the decoupled / single-robot solve.
  double* cpl_part;   // robots x 7 : a.y, a.z, g.g, h_t, g_t, z.g, y.g
  double* cpl_zyg;    // robots x (N*9) x 3 : z, y, g
  DevCounts* dc;      // a failed factorisation sets TOB_ERR_SOLVE: the iteration is not committed
};

__global__ void __launch_bounds__(1024) k_solve_bcr(BcrArgs a) {
  extern __shared__ double sm[];
  const int robot = a.robot_begin + blockIdx.x;
  const int N = a.P + 1, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, nw = blockDim.x >> 5;
  double* blk = sm;                               // N x BLK
  double* rhs0 = blk + (size_t)N * BLK;           // N x 18 : original [g | arrow] per block (for the Schur step)
  __shared__ double s_h, s_gt, s_t;
  __shared__ int s_fail;
  __shared__ double s_red[32][4];
  if (tid == 0) s_fail = 0;
  const double* G = a.pc_g + (size_t)19 * robot * a.P;
  const double* H = a.pc_h + (size_t)361 * robot * a.P;
  // ---- assembly
  for (int e = tid; e < N * 81; e += blockDim.x) {
    const int b = e / 81, rc = e - 81 * b, r = rc % BS, cc = rc / BS;
    double d = 0, el = 0, eu = 0;
    if (b >= 1) { const double* h = H + (size_t)361 * (b - 1); d += h[(9 + r) + 19 * (9 + cc)]; el = h[(9 + r) + 19 * cc]; }
    if (b < a.P) { const double* h = H + (size_t)361 * b; d += h[r + 19 * cc]; eu = h[r + 19 * (9 + cc)]; }
    // fixed end control points: block 0 local 0..5, block N-1 local 3..8
    const bool fr = (b == 0 && r < 6) || (b == N - 1 && r >= 3);
    const bool fc = (b == 0 && cc < 6) || (b == N - 1 && cc >= 3);
    if (fr || fc) d = (r == cc) ? 1.0 : 0.0;
    if (fr || (b == 1 && cc < 6)) el = 0.0;            // A(1,0) columns of fixed unknowns
    if (fr || (b == N - 2 && cc >= 3)) eu = 0.0;       // A(N-2,N-1) columns of fixed unknowns
    double* o = blk + (size_t)b * BLK;
    o[rc] = el; o[81 + rc] = d; o[162 + rc] = eu;
  }
  for (int e = tid; e < N * 18; e += blockDim.x) {
    const int b = e / 18, q = e - 18 * b, r = q % BS, col = q / BS;
    double v = 0;
    if (b >= 1) v += col == 0 ? G[(size_t)19 * (b - 1) + 9 + r] : H[(size_t)361 * (b - 1) + (9 + r) + 19 * 18];
    if (b < a.P) v += col == 0 ? G[(size_t)19 * b + r] : H[(size_t)361 * b + r + 19 * 18];
    if ((b == 0 && r < 6) || (b == N - 1 && r >= 3)) v = 0.0;
    blk[(size_t)b * BLK + 243 + q] = v;
    rhs0[e] = v;
  }
  if (tid == 0) {
    double h = 0, gt = 0;
    for (int sp = 0; sp < a.P; sp++) { h += H[(size_t)361 * sp + 18 + 19 * 18]; gt += G[(size_t)19 * sp + 18]; }
    s_h = h; s_gt = gt;
  }
  __syncthreads();
  // ---- reduction
  int s = 1;
  for (; s < N; s <<= 1) {
    // eliminate blocks i = s, 3s, 5s, ...
    const int n_el = (N - 1 - s) / (2 * s) + 1;
    for (int q = wp; q < n_el; q += nw) {
      const int i = s + 2 * s * q;
      double* B = blk + (size_t)i * BLK;
      double *El = B, *D = B + 81, *Eu = B + 162, *R = B + 243, *dinv = B + 261;
      const bool has_up = i + s < N;
      if (!warp_chol9(D, dinv)) { if (lane == 0) s_fail = 1; }
      // forward substitution L y = b for the 20 columns [El | Eu | R]
      if (lane < 20) {
        double* col = lane < 9 ? El + BS * lane : (lane < 18 ? Eu + BS * (lane - 9) : R + BS * (lane - 18));
        double y[BS];
        if (lane >= 9 && lane < 18 && !has_up) {
#pragma unroll
          for (int k = 0; k < BS; k++) col[k] = 0.0;
        } else {
          // right-looking: once y[k] is known the remaining rows are updated by independent multiply-adds, so the
          // dependency chain is 9 x (mul + fma) instead of the 45 chained fma of the dot-product form
#pragma unroll
          for (int k = 0; k < BS; k++) y[k] = col[k];
#pragma unroll
          for (int k = 0; k < BS; k++) {
            y[k] *= dinv[k];
#pragma unroll
            for (int i = k + 1; i < BS; i++) y[i] -= D[i + BS * k] * y[k];
          }
#pragma unroll
          for (int k = 0; k < BS; k++) col[k] = y[k];
        }
      }
    }
    __syncthreads();
    // update survivors j = 0, 2s, 4s, ...
    const int n_sv = (N - 1) / (2 * s) + 1;
    for (int q = wp; q < n_sv; q += nw) {
      const int j = 2 * s * q;
      double* B = blk + (size_t)j * BLK;
      const double* lo = (j - s >= 0) ? blk + (size_t)(j - s) * BLK : nullptr;        // eliminated lower neighbour: uses its V, U, c
      const double* up = (j + s < N) ? blk + (size_t)(j + s) * BLK : nullptr;         // eliminated upper neighbour: uses its U, V, c
      for (int e = lane; e < 81; e += 32) {
        const int r = e % BS, cc = e / BS;
        double d = 0, nel = 0, neu = 0;
        if (lo) {
          const double *U = lo, *V = lo + 162;
#pragma unroll
          for (int k = 0; k < BS; k++) { d += V[k + BS * r] * V[k + BS * cc]; nel -= V[k + BS * r] * U[k + BS * cc]; }
        }
        if (up) {
          const double *U = up, *V = up + 162;
#pragma unroll
          for (int k = 0; k < BS; k++) { d += U[k + BS * r] * U[k + BS * cc]; neu -= U[k + BS * r] * V[k + BS * cc]; }
        }
        B[81 + e] -= d;
        B[e] = nel;          // A(j, j-2s)
        B[162 + e] = neu;    // A(j, j+2s)
      }
      if (lane < 18) {
        const int r = lane % BS, col = lane / BS;
        double d = 0;
        if (lo) {
          const double *V = lo + 162, *cv = lo + 243 + BS * col;
#pragma unroll
          for (int k = 0; k < BS; k++) d += V[k + BS * r] * cv[k];
        }
        if (up) {
          const double *U = up, *cv = up + 243 + BS * col;
#pragma unroll
          for (int k = 0; k < BS; k++) d += U[k + BS * r] * cv[k];
        }
        B[243 + lane] -= d;
      }
    }
    __syncthreads();
  }
  // ---- root block 0
  if (wp == 0) {
    double* B = blk;
    double *D = B + 81, *R = B + 243, *dinv = B + 261;
    if (!warp_chol9(D, dinv)) { if (lane == 0) s_fail = 1; }
    if (lane < 2) {
      double* col = R + BS * lane;
      double y[BS];
#pragma unroll
      for (int k = 0; k < BS; k++) y[k] = col[k];
#pragma unroll
      for (int k = 0; k < BS; k++) {
        y[k] *= dinv[k];
#pragma unroll
        for (int i = k + 1; i < BS; i++) y[i] -= D[i + BS * k] * y[k];
      }
#pragma unroll
      for (int k = BS - 1; k >= 0; k--) {
        y[k] *= dinv[k];
#pragma unroll
        for (int i = 0; i < k; i++) y[i] -= D[k + BS * i] * y[k];
      }
#pragma unroll
      for (int k = 0; k < BS; k++) col[k] = y[k];
    }
  }
  __syncthreads();
  // ---- back substitution, levels in reverse
  for (s >>= 1; s >= 1; s >>= 1) {
    const int n_el = (N - 1 - s) / (2 * s) + 1;
    for (int q = wp; q < n_el; q += nw) {
      const int i = s + 2 * s * q;
      double* B = blk + (size_t)i * BLK;
      const double *U = B, *D = B + 81, *V = B + 162, *dinv = B + 261;
      double* R = B + 243;
      const double* xl = blk + (size_t)(i - s) * BLK + 243;
      const double* xu = (i + s < N) ? blk + (size_t)(i + s) * BLK + 243 : nullptr;
      double t = 0;
      if (lane < 18) {
        const int r = lane % BS, col = lane / BS;
        t = R[lane];
#pragma unroll
        for (int k = 0; k < BS; k++) t -= U[r + BS * k] * xl[k + BS * col];
        if (xu) {
#pragma unroll
          for (int k = 0; k < BS; k++) t -= V[r + BS * k] * xu[k + BS * col];
        }
      }
      __syncwarp();
      if (lane < 18) R[lane] = t;
      __syncwarp();
      if (lane < 2) {
        double* col = R + BS * lane;
        double y[BS];
#pragma unroll
        for (int k = 0; k < BS; k++) y[k] = col[k];
#pragma unroll
        for (int k = BS - 1; k >= 0; k--) {
          y[k] *= dinv[k];
#pragma unroll
          for (int i = 0; i < k; i++) y[i] -= D[k + BS * i] * y[k];
        }
#pragma unroll
        for (int k = 0; k < BS; k++) col[k] = y[k];
      }
    }
    __syncthreads();
  }
  // ---- Schur complement for the arrow (shared piece time), direction, wolfe, gnorm.  z = B^-1 g (col 0), y = B^-1 a (col 1)
  double ay = 0, az = 0, gg = 0;
  for (int e = tid; e < N * BS; e += blockDim.x) {
    const int b = e / BS, r = e - BS * b;
    const double g0 = rhs0[b * 18 + r], a0 = rhs0[b * 18 + BS + r];
    const double z = blk[(size_t)b * BLK + 243 + r], y = blk[(size_t)b * BLK + 243 + BS + r];
    ay += a0 * y; az += a0 * z; gg += g0 * g0;
  }
  for (int o = 16; o; o >>= 1) {
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
    az += __shfl_xor_sync(0xffffffffu, az, o);
    gg += __shfl_xor_sync(0xffffffffu, gg, o);
  }
  if (lane == 0) { s_red[wp][0] = ay; s_red[wp][1] = az; s_red[wp][2] = gg; }
  __syncthreads();
  if (a.cpl_part) {
    // coupled system: this robot's contributions to the shared-time Schur complement and to wolfe = -(x.g + t g_t) with
    // x_i = -z_i - y_i t, i.e. x_i.g_i = -(z_i.g_i) - t (y_i.g_i): seven sums per robot, closed over ALL robots (in robot
    // order, whichever rank computed them) by k_couple_finish
    double zg = 0, yg = 0;
    for (int e = tid; e < N * BS; e += blockDim.x) {
      const int b = e / BS, r = e - BS * b;
      const double g0 = rhs0[b * 18 + r];
      zg += blk[(size_t)b * BLK + 243 + r] * g0;
      yg += blk[(size_t)b * BLK + 243 + BS + r] * g0;
    }
    for (int o = 16; o; o >>= 1) { zg += __shfl_xor_sync(0xffffffffu, zg, o); yg += __shfl_xor_sync(0xffffffffu, yg, o); }
    __shared__ double s_red2[32][2];
    if (lane == 0) { s_red2[wp][0] = zg; s_red2[wp][1] = yg; }
    __syncthreads();
    if (tid == 0) {
      double AY = 0, AZ = 0, GG = 0, ZG = 0, YG = 0;
      for (int i = 0; i < nw; i++) { AY += s_red[i][0]; AZ += s_red[i][1]; GG += s_red[i][2]; ZG += s_red2[i][0]; YG += s_red2[i][1]; }
      double* o = a.cpl_part + 7 * (size_t)robot;
      o[0] = AY; o[1] = AZ; o[2] = GG; o[3] = s_h; o[4] = s_gt; o[5] = ZG; o[6] = YG;
      a.status[robot] = s_fail;
      if (s_fail && a.dc && !(a.dc->overflow & TOB_OVF_RETRY)) atomicOr(&a.dc->overflow, TOB_ERR_SOLVE);
    }
    double* o = a.cpl_zyg + (size_t)robot * N * BS * 2;
    for (int e = tid; e < N * BS; e += blockDim.x) {
      const int b = e / BS, r = e - BS * b;
      o[2 * e] = blk[(size_t)b * BLK + 243 + r];
      o[2 * e + 1] = blk[(size_t)b * BLK + 243 + BS + r];
    }
    return;
  }
  if (tid == 0) {
    double AY = 0, AZ = 0, GG = 0;
    for (int i = 0; i < nw; i++) { AY += s_red[i][0]; AZ += s_red[i][1]; GG += s_red[i][2]; }
    const double schur = s_h - AY;
    if (!(schur > 0)) s_fail = 1;
    s_t = (AZ - s_gt) / schur;
    a.tdir[robot] = s_t;
    a.gnorm[robot] = sqrt(GG + s_gt * s_gt);
  }
  __syncthreads();
  const double t = s_t;
  double wl = 0;
  double* dir = a.dir + (size_t)robot * 3 * a.T;
  for (int e = tid; e < N * BS; e += blockDim.x) {       // full coordinate f = e: control point f/3, axis f%3
    const int b = e / BS, r = e - BS * b;
    const double z = blk[(size_t)b * BLK + 243 + r], y = blk[(size_t)b * BLK + 243 + BS + r];
    const bool fixed = (b == 0 && r < 6) || (b == N - 1 && r >= 3);
    const double x = fixed ? 0.0 : (-z - y * t);
    wl += x * rhs0[b * 18 + r];
    dir[(size_t)(e % 3) * a.T + e / 3] = x;
  }
  for (int o = 16; o; o >>= 1) wl += __shfl_xor_sync(0xffffffffu, wl, o);
  if (lane == 0) s_red[wp][3] = wl;
  __syncthreads();
  if (tid == 0) {
    double W = 0;
    for (int i = 0; i < nw; i++) W += s_red[i][3];
    W += t * s_gt;
    a.wolfe[robot] = -W;
    a.status[robot] = s_fail;
    if (s_fail && a.dc && !(a.dc->overflow & TOB_OVF_RETRY)) atomicOr(&a.dc->overflow, TOB_ERR_SOLVE);
  }
}

// closes the shared-time Schur complement of the coupled system: t = (sum a_i.z_i - sum g_t,i) / (sum h_t,i - sum a_i.y_i),
// x_i = -z_i - y_i t, wolfe = -(sum x_i.g_i + t sum g_t,i), gnorm = |G| / U (Optimization3D_multi.h:553-583).  `part` holds
// the seven sums of EVERY robot (sharded: all-gathered), summed in robot order, so every rank gets bitwise the same t, wolfe
// and gnorm as a single context; the directions are written for the owned robots [rb, re).
__global__ void __launch_bounds__(256) k_couple_finish(const double* __restrict__ part, const double* __restrict__ zy, int U, int rb,
                                                       int re, int N, int T, double* dir, double* tdir, double* wolfe, double* gnorm,
                                                       DevCounts* dc) {
  __shared__ double s_t;
  const int tid = threadIdx.x;
  if (tid == 0) {
    double AY = 0, AZ = 0, GG = 0, HT = 0, GT = 0, ZG = 0, YG = 0;
    for (int u = 0; u < U; u++) {
      const double* p = part + 7 * (size_t)u;
      AY += p[0]; AZ += p[1]; GG += p[2]; HT += p[3]; GT += p[4]; ZG += p[5]; YG += p[6];
    }
    const double schur = HT - AY;
    if (!(schur > 0) && !(dc->overflow & TOB_OVF_RETRY)) atomicOr(&dc->overflow, TOB_ERR_SOLVE);
    const double t = (AZ - GT) / schur;
    s_t = t;
    const double gn = sqrt(GG + GT * GT) / double(U);
    const double w = (ZG + t * YG) - t * GT;
    for (int u = 0; u < U; u++) { tdir[u] = t; gnorm[u] = gn; wolfe[u] = w; }
  }
  __syncthreads();
  const double t = s_t;
  const int per = N * BS;
  for (int i = tid; i < (re - rb) * per; i += blockDim.x) {
    const int u = rb + i / per, e = i % per, b = e / BS, r = e - BS * b;
    const size_t g = (size_t)u * per + e;
    const bool fixed = (b == 0 && r < 6) || (b == N - 1 && r >= 3);
    const double x = fixed ? 0.0 : (-zy[2 * g] - zy[2 * g + 1] * t);
    dir[(size_t)u * 3 * T + (size_t)(e % 3) * T + e / 3] = x;
  }
}

// coupled Newton direction of the owned robots.  Sharded: the per-robot Schur sums are exchanged between the two kernels.
int solve_coupled(tob_ctx* c) {
  const int U = c->n_robots(), N = c->prm.piece_num + 1, rb = c->own_begin, re = c->own_end;
  const size_t smem_bcr = ((size_t)N * BLK + (size_t)N * 18) * sizeof(double);
  if (smem_bcr > 220 * 1024) return fail_msg(c, "coupled mode: trajectories with more than ~100 pieces are not supported");
  TOB_CUDA(c, c->cpl_part.ensure((size_t)7 * U));
  TOB_CUDA(c, c->cpl_zy.ensure((size_t)2 * U * N * BS));
  BcrArgs a;
  a.pc_g = c->pc_g.p; a.pc_h = c->pc_h.p; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb;
  a.dir = c->s_dir.p; a.tdir = c->s_tdir.p; a.wolfe = c->s_wolfe.p; a.gnorm = c->s_gnorm.p; a.status = c->solve_status.p;
  a.cpl_part = c->cpl_part.p; a.cpl_zyg = c->cpl_zy.p; a.dc = c->dc.p;
  if (!c->bcr_attr_set) {
    TOB_CUDA(c, cudaFuncSetAttribute(k_solve_bcr, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    c->bcr_attr_set = true;
  }
  int warps = (N + 1) / 2;
  if (warps < 4) warps = 4;
  if (warps > 32) warps = 32;
  {
    Prof prof(c, K_SOLVE);
    k_solve_bcr<<<re - rb, warps * 32, smem_bcr, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
  }
  TOB_TRY(exchange_robots(c, c->cpl_part.p, 7, sizeof(double)));
  k_couple_finish<<<1, 256, 0, c->stream>>>(c->cpl_part.p, c->cpl_zy.p, U, rb, re, N, c->T, c->s_dir.p, c->s_tdir.p, c->s_wolfe.p,
                                            c->s_gnorm.p, c->dc.p);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

int solve_directions(tob_ctx* c, int rb, int re, int dense_shift) {
  int nr = c->n_robots();
  TOB_CUDA(c, c->s_dir.ensure((size_t)3 * c->T * nr));
  TOB_CUDA(c, c->s_tdir.ensure(nr));
  TOB_CUDA(c, c->s_wolfe.ensure(nr));
  TOB_CUDA(c, c->s_gnorm.ensure(nr));
  TOB_CUDA(c, c->solve_status.ensure(nr));
  const int N = c->prm.piece_num + 1;
  const size_t smem_bcr = ((size_t)N * BLK + (size_t)N * 18) * sizeof(double);
  if (smem_bcr <= 220 * 1024) {
    BcrArgs a;
    a.pc_g = c->pc_g.p; a.pc_h = c->pc_h.p; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb;
    a.dir = c->s_dir.p; a.tdir = c->s_tdir.p; a.wolfe = c->s_wolfe.p; a.gnorm = c->s_gnorm.p; a.status = c->solve_status.p;
    a.cpl_part = nullptr; a.cpl_zyg = nullptr; a.dc = c->dc.p;
    if (!c->bcr_attr_set) {
      TOB_CUDA(c, cudaFuncSetAttribute(k_solve_bcr, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      c->bcr_attr_set = true;
    }
    int warps = (N + 1) / 2;
    if (warps < 4) warps = 4;
    if (warps > 32) warps = 32;
    Prof prof(c, K_SOLVE);
    k_solve_bcr<<<re - rb, warps * 32, smem_bcr, c->stream>>>(a);
    TOB_LAUNCH_CHECK(c);
    return 0;
  }
  // very long trajectories (P > ~100): sequential banded factorisation with a global-memory band
  SolveArgs a;
  a.pc_g = c->pc_g.p; a.pc_h = c->pc_h.p; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb; a.dense_shift = dense_shift;
  a.dir = c->s_dir.p; a.tdir = c->s_tdir.p; a.wolfe = c->s_wolfe.p; a.gnorm = c->s_gnorm.p; a.status = c->solve_status.p;
  int m = 3 * (c->T - 4);
  TOB_CUDA(c, c->band.ensure((size_t)(re - rb) * m * 22));
  a.gband = c->band.p;
  a.use_global = 1;
  a.dc = c->dc.p;
  Prof prof(c, K_SOLVE);
  k_solve<<<re - rb, 256, 0, c->stream>>>(a);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// ---- slack / dual update ---------------------------------------------------------------------------------------------
// One WARP per (robot, piece); the 19x19 system lives in shared memory (lane i owns row i in the Cholesky and the
// substitutions).  H = (ks/t^5) M_dynamic (x) I3 + mu I with one dense arrow row/column for the piece time; the reference
// factors it with Eigen::LLT and falls back to an eigenvalue shift when that fails (Optimization3D_admm.h:313-327).
struct SlackArgs {
  const double *spline, *ptime, *convert, *mdyn;
  double *pslack, *tslack, *plambda, *tlambda;
  double mu, ks, kt;
  int P, T, robot_begin, n;
  DevCounts* guard;   // non-null: skip when the iteration cannot be committed, else count it in guard->iters_done
};

struct SlackWarp {       // per-warp shared scratch
  double H[361];         // Gradient_admm::slack_gradient hessian, col-major ld 19
  double A[361];         // reduced system -> Cholesky factor
  double g[19], b[19], x[19];
  double cs[18], p[18], lam[18], dir[18], pn[18];   // [m][k] flatten (row-major 6x3), as the reference's transposeInPlace + Map
  double ws[4 * 19];
};

__device__ __forceinline__ double wsum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Gradient_admm::dynamic_gradient :633-671 (consensus == 0: g[18] = g_t, H(18,18) = h_t, arrow = partgrad) or
// Gradient_admm::slack_gradient :574-622 (consensus == 1) of the point S.p / t.  Returns the jerk energy.
__device__ double slack_build_warp(SlackWarp& S, const double* __restrict__ M, double ptime, double t, double tlam, double mu,
                                   double ks, double kt, int consensus) {
  const int lane = threadIdx.x & 31;
  const double c5 = ks / pow(t, 5.0);
  double q = 0, gdyn = 0;
  if (lane < 18) {
    const int r = lane / 3, k = lane - 3 * r;
    double mx = 0;
#pragma unroll
    for (int s2 = 0; s2 < 6; s2++) mx += M[r + 6 * s2] * S.p[s2 * 3 + k];
    gdyn = c5 * mx;
    q = S.p[lane] * mx;
  }
  const double dyn = c5 * 0.5 * wsum(q);
  for (int e = lane; e < 361; e += 32) {
    const int r = e % 19, cc = e / 19;
    double v = 0;
    if (r < 18 && cc < 18) {
      if (r % 3 == cc % 3) v = c5 * M[r / 3 + 6 * (cc / 3)];
      if (consensus && r == cc) v += mu;
    }
    S.H[e] = v;
  }
  __syncwarp();
  if (lane < 18) {
    const double pg = -5 * gdyn / t;
    S.H[lane + 19 * 18] = pg;
    S.H[18 + 19 * lane] = pg;
    S.g[lane] = consensus ? gdyn + (mu * (S.p[lane] - S.cs[lane]) - S.lam[lane]) : gdyn;
  }
  if (lane == 0) {
    double g_t = -5 * dyn / t + kt * 1.1 * pow(t, 0.1);
    double h_t = 30 * dyn / (t * t) + kt * 0.11 * pow(t, -0.9);
    if (consensus) { g_t += mu * (t - ptime) - tlam; h_t += mu; }
    S.g[18] = g_t;
    S.H[18 + 19 * 18] = h_t;
  }
  __syncwarp();
  return dyn;
}

// Energy_admm::slack_energy :172-190 (consensus == 1) / dynamic_energy :199-215 (consensus == 0) of the point pv / t
__device__ double slack_energy_warp(const SlackWarp& S, const double* pv, const double* __restrict__ M, double ptime, double t,
                                    double tlam, double mu, double ks, double kt, int consensus) {
  const int lane = threadIdx.x & 31;
  double q = 0, sq = 0, lin = 0;
  if (lane < 18) {
    const int r = lane / 3, k = lane - 3 * r;
    double mx = 0;
#pragma unroll
    for (int s2 = 0; s2 < 6; s2++) mx += M[r + 6 * s2] * pv[s2 * 3 + k];
    q = pv[lane] * mx;
    const double dlt = S.cs[lane] - pv[lane];
    sq = dlt * dlt;
    lin = S.lam[lane] * dlt;
  }
  q = wsum(q); sq = wsum(sq); lin = wsum(lin);
  double e = ks / pow(t, 5.0) * 0.5 * q + kt * pow(t, 1.1);
  if (consensus) {
    e += mu / 2.0 * sq;
    e += mu / 2.0 * (ptime - t) * (ptime - t);
    e += lin;
    e += tlam * (ptime - t);
  }
  return e;
}

// L L^T x = b by the warp (L col-major ld n in its lower triangle, lane i owns row i); v (shared) in: b, out: x
__device__ void warp_chol_solve(const double* L, int n, double* v) {
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < n; k++) {
    if (lane == k) v[k] = v[k] / L[k + n * k];
    __syncwarp();
    if (lane > k && lane < n) v[lane] -= L[lane + n * k] * v[k];
    __syncwarp();
  }
  for (int k = n - 1; k >= 0; k--) {
    if (lane == k) v[k] = v[k] / L[k + n * k];
    __syncwarp();
    if (lane < k) v[lane] -= L[k + n * lane] * v[k];
    __syncwarp();
  }
}

// reduced Newton system of one piece (N = 19, or 13 at the two ends): Cholesky with the reference's eigenvalue-shift fall-back
// (Optimization3D_admm.h:313-327), then x = H^-1 g.  S.A holds the reduced matrix (col-major, ld N) on entry and L on exit.
template <int N>
__device__ __forceinline__ void slack_newton_solve(SlackWarp& S) {
  const int lane = threadIdx.x & 31;
  double r[N];
#pragma unroll
  for (int j = 0; j < N; j++) r[j] = lane < N ? S.A[lane + N * j] : 0.0;
  const double b = lane < N ? S.x[lane] : 0.0;
  double* L = S.H;                         // the full 19x19 Hessian has been consumed: its storage takes the factor
  if (!warp_chol_roll<N>(r, L)) {
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = lane < N ? S.A[lane + N * j] : 0.0;
    const double mn = warp_min_eig_roll<N>(r, S.ws, S.ws + 19);
#pragma unroll
    for (int j = 0; j < N; j++) {
      double v = lane < N ? S.A[lane + N * j] : 0.0;
      if (lane == j && mn < 0) v = v - mn * 1.0 + 0.01 * 1.0;
      r[j] = v;
    }
    __syncwarp();
    warp_chol_roll<N>(r, L);
  }
  __syncwarp();
  const double x = warp_chol_solve_sm<N>(L, b);
  if (lane < N) S.x[lane] = x;
  __syncwarp();
}

#define SLACK_WARPS 4
__global__ void __launch_bounds__(32 * SLACK_WARPS) k_slack(SlackArgs a) {
  __shared__ SlackWarp sm[SLACK_WARPS];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int idx = blockIdx.x * SLACK_WARPS + wp;
  if (iteration_blocked(a.guard)) return;                     // uniform over the grid: nothing in this kernel changes it
  if (a.guard && blockIdx.x == 0 && threadIdx.x == 0) a.guard->iters_done++;
  if (idx >= a.n) return;                                     // whole warp leaves together
  SlackWarp& S = sm[wp];
  const int robot = a.robot_begin + idx / a.P, sp = idx % a.P;
  const double* C = a.convert + (size_t)36 * sp;
  const double* M = a.mdyn;
  const double ptime = a.ptime[robot];
  if (lane < 18) {
    const int r = lane / 3, k = lane - 3 * r;
    double acc = 0;
#pragma unroll
    for (int kk = 0; kk < 6; kk++) acc += C[r + 6 * kk] * a.spline[(size_t)robot * 3 * a.T + (size_t)k * a.T + 3 * sp + kk];
    S.cs[lane] = acc;
    const size_t s = (size_t)robot * 18 * a.P + (size_t)k * 6 * a.P + 6 * sp + r;
    S.p[lane] = a.pslack[s];
    S.lam[lane] = a.plambda[s];
    S.dir[lane] = 0.0;
  }
  const size_t pb = (size_t)robot * a.P + sp;
  const double t0 = a.tslack[pb], tlam = a.tlambda[pb];
  __syncwarp();
  slack_build_warp(S, M, ptime, t0, tlam, a.mu, a.ks, a.kt, 1);
  // reduced system: first piece drops control points 0,1; last piece drops control points 4,5
  int off = 0, tn = 6;
  if (sp == 0) { off = 6; tn = 4; }
  else if (sp == a.P - 1) { off = 0; tn = 4; }
  const int n = 3 * tn + 1;
#define GI(i) ((i) < 3 * tn ? off + (i) : 18)
  for (int e = lane; e < n * n; e += 32) { const int i = e % n, j = e / n; S.A[e] = S.H[GI(i) + 19 * GI(j)]; }
  if (lane < n) { S.b[lane] = S.g[GI(lane)]; S.x[lane] = S.b[lane]; }
  __syncwarp();
#undef GI
  if (n == 19) slack_newton_solve<19>(S);
  else slack_newton_solve<13>(S);                               // x = H^-1 g ; the Newton step is -x
  double wl = (lane < n) ? S.x[lane] * S.b[lane] : 0.0;
  const double wolfe = wsum(wl);                               // = -(-x).g
  if (lane < 3 * tn) S.dir[(sp == 0 ? 6 : 0) + lane] = -S.x[lane];
  __syncwarp();
  const double tdir = -S.x[3 * tn];
  double step = 1.0;
  if (t0 + step * tdir <= 0) step = -0.95 * t0 / tdir;
  const double e0 = slack_energy_warp(S, S.p, M, ptime, t0, tlam, a.mu, a.ks, a.kt, 1);
  double t = t0 + step * tdir;
  for (int guard = 0; guard < 400; guard++) {
    if (lane < 18) S.pn[lane] = S.p[lane] + step * S.dir[lane];
    __syncwarp();
    const double e1 = slack_energy_warp(S, S.pn, M, ptime, t, tlam, a.mu, a.ks, a.kt, 1);
    if (!(e0 - 1e-4 * wolfe * step < e1)) break;               // uniform: e1 is identical on every lane
    step *= 0.8;
    t = t0 + step * tdir;
    __syncwarp();
  }
  if (lane < 18) {
    const int r = lane / 3, k = lane - 3 * r;
    const double pn = S.p[lane] + step * S.dir[lane];
    const size_t s = (size_t)robot * 18 * a.P + (size_t)k * 6 * a.P + 6 * sp + r;
    a.pslack[s] = pn;
    a.plambda[s] = S.lam[lane] + a.mu * (S.cs[lane] - pn);
  }
  if (lane == 0) {
    a.tslack[pb] = t;
    a.tlambda[pb] = tlam + a.mu * (ptime - t);
  }
}

int slack_update(tob_ctx* c, int rb, int re, int guarded) {
  SlackArgs a;
  a.guard = guarded ? c->dc.p : nullptr;
  a.spline = c->s_spline.p; a.ptime = c->s_ptime.p; a.convert = c->d_convert.p; a.mdyn = c->d_mdyn.p;
  a.pslack = c->s_pslack.p; a.tslack = c->s_tslack.p; a.plambda = c->s_plambda.p; a.tlambda = c->s_tlambda.p;
  a.mu = c->prm.mu; a.ks = c->prm.ks; a.kt = c->prm.kt; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb;
  a.n = (re - rb) * c->prm.piece_num;
  Prof prof(c, K_SLACK);
  k_slack<<<div_up(a.n, SLACK_WARPS), 32 * SLACK_WARPS, 0, c->stream>>>(a);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// function-level entry point: one piece, one warp.  in: cs[18] p[18] lam[18] (6x3 col-major each), scal = {ptime, t, tlam}
// out: energy, g[19], H[361]
__global__ void k_slack_terms(const double* in, const double* mdyn, double mu, double ks, double kt, int consensus, double* out) {
  __shared__ SlackWarp S;
  const int lane = threadIdx.x & 31;
  if (lane < 18) {
    const int r = lane / 3, k = lane - 3 * r;          // [m][k] <- col-major 6x3
    S.cs[lane] = in[r + 6 * k];
    S.p[lane] = in[18 + r + 6 * k];
    S.lam[lane] = in[36 + r + 6 * k];
  }
  __syncwarp();
  const double ptime = in[54], t = in[55], tlam = in[56];
  const double e = slack_energy_warp(S, S.p, mdyn, ptime, t, tlam, mu, ks, kt, consensus);
  slack_build_warp(S, mdyn, ptime, t, tlam, mu, ks, kt, consensus);
  if (lane == 0) out[0] = e;
  if (lane < 19) out[1 + lane] = S.g[lane];
  for (int i = lane; i < 361; i += 32) out[20 + i] = S.H[i];
}

int slack_terms(tob_ctx* c, const double* in57_dev, int consensus, double* out381_dev) {
  k_slack_terms<<<1, 32, 0, c->stream>>>(in57_dev, c->d_mdyn.p, c->prm.mu, c->prm.ks, c->prm.kt, consensus, out381_dev);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

}  // namespace tob
