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
  }
}

int solve_directions(tob_ctx* c, int rb, int re, int dense_shift) {
  int nr = c->n_robots();
  TOB_CUDA(c, c->s_dir.ensure((size_t)3 * c->T * nr));
  TOB_CUDA(c, c->s_tdir.ensure(nr));
  TOB_CUDA(c, c->s_wolfe.ensure(nr));
  TOB_CUDA(c, c->s_gnorm.ensure(nr));
  TOB_CUDA(c, c->solve_status.ensure(nr));
  SolveArgs a;
  a.pc_g = c->pc_g.p; a.pc_h = c->pc_h.p; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb; a.dense_shift = dense_shift;
  a.dir = c->s_dir.p; a.tdir = c->s_tdir.p; a.wolfe = c->s_wolfe.p; a.gnorm = c->s_gnorm.p; a.status = c->solve_status.p;
  int m = 3 * (c->T - 4);
  size_t smem = (size_t)m * 22 * sizeof(double);
  a.use_global = smem > 200 * 1024;
  a.gband = nullptr;
  if (a.use_global) {
    TOB_CUDA(c, c->band.ensure((size_t)(re - rb) * m * 22));
    a.gband = c->band.p;
    smem = 0;
  } else if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      TOB_CUDA(c, cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
  }
  Prof prof(c, K_SOLVE);
  k_solve<<<re - rb, 256, smem, c->stream>>>(a);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

// ---- slack / dual update ---------------------------------------------------------------------------------------------
struct SlackArgs {
  const double *spline, *ptime, *convert, *mdyn;
  double *pslack, *tslack, *plambda, *tlambda;
  double mu, ks, kt;
  int P, T, robot_begin, n;
};

__device__ double dyn_energy(const double* M, const double* p /*[6][3] as p[m*3+k]*/, double t, double ks, double kt) {
  double e = 0;
  double c5 = ks / pow(t, 5.0);
  for (int k = 0; k < 3; k++) {
    double q = 0;
    for (int r = 0; r < 6; r++) {
      double mx = 0;
      for (int s = 0; s < 6; s++) mx += M[r + 6 * s] * p[s * 3 + k];
      q += p[r * 3 + k] * mx;
    }
    e += c5 * 0.5 * q;
  }
  return e;
}

__device__ double slack_energy_dev(const double* M, const double* cs, double ptime, const double* p, double t, const double* lam,
                                   double tlam, double mu, double ks, double kt) {
  double e = dyn_energy(M, p, t, ks, kt) + kt * pow(t, 1.1);
  double sq = 0, lin = 0;
  for (int i = 0; i < 18; i++) {
    double dlt = cs[i] - p[i];
    sq += dlt * dlt;
    lin += lam[i] * dlt;
  }
  e += mu / 2.0 * sq;
  e += mu / 2.0 * (ptime - t) * (ptime - t);
  e += lin;
  e += tlam * (ptime - t);
  return e;
}

__global__ void __launch_bounds__(32) k_slack(SlackArgs a) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n) return;
  const int robot = a.robot_begin + idx / a.P, sp = idx % a.P;
  const double* C = a.convert + (size_t)36 * sp;
  const double* M = a.mdyn;
  const double ptime = a.ptime[robot];
  // local copies, layout [m][k] (row-major flatten of the 6x3 blocks, as the reference's transposeInPlace + Map)
  double cs[18], p[18], lam[18];
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 6; r++) {
      double acc = 0;
      for (int kk = 0; kk < 6; kk++) acc += C[r + 6 * kk] * a.spline[(size_t)robot * 3 * a.T + (size_t)k * a.T + 3 * sp + kk];
      cs[r * 3 + k] = acc;
      size_t s = (size_t)robot * 18 * a.P + (size_t)k * 6 * a.P + 6 * sp + r;
      p[r * 3 + k] = a.pslack[s];
      lam[r * 3 + k] = a.plambda[s];
    }
  const size_t pb = (size_t)robot * a.P + sp;
  double t = a.tslack[pb], tlam = a.tlambda[pb];
  // gradient / Hessian (19)
  double g[19], H[19 * 19];
  for (int i = 0; i < 361; i++) H[i] = 0;
  const double c5 = a.ks / pow(t, 5.0);
  double dyn = 0;
  for (int k = 0; k < 3; k++) {
    double q = 0;
    for (int r = 0; r < 6; r++) {
      double mx = 0;
      for (int s = 0; s < 6; s++) mx += M[r + 6 * s] * p[s * 3 + k];
      g[r * 3 + k] = c5 * mx;
      q += p[r * 3 + k] * mx;
    }
    dyn += c5 * 0.5 * q;
  }
  for (int r = 0; r < 6; r++)
    for (int s = 0; s < 6; s++)
      for (int k = 0; k < 3; k++) H[(r * 3 + k) + 19 * (s * 3 + k)] = c5 * M[r + 6 * s];
  double g_t = -5 * dyn / t + a.kt * 1.1 * pow(t, 0.1);
  double h_t = 30 * dyn / (t * t) + a.kt * 0.11 * pow(t, -0.9);
  for (int i = 0; i < 18; i++) {
    double pg = -5 * g[i] / t;
    H[i + 19 * 18] = pg;
    H[18 + 19 * i] = pg;
    g[i] += a.mu * (p[i] - cs[i]) - lam[i];
    H[i + 19 * i] += a.mu;
  }
  g_t += a.mu * (t - ptime) - tlam;
  h_t += a.mu;
  g[18] = g_t;
  H[18 + 19 * 18] = h_t;
  // reduced system: first piece drops control points 0,1; last piece drops control points 4,5
  int off = 0, tn = 6;
  if (sp == 0) { off = 6; tn = 4; }
  else if (sp == a.P - 1) { off = 0; tn = 4; }
  const int n = 3 * tn + 1;
  double A[19 * 19], L[19 * 19], b[19], x[19];
  auto gi = [&](int i) { return i < 3 * tn ? off + i : 18; };
  for (int i = 0; i < n; i++) {
    b[i] = g[gi(i)];
    for (int j = 0; j < n; j++) A[i + n * j] = H[gi(i) + 19 * gi(j)];
  }
  if (!chol_is_spd_n(A, L, n)) {
    for (int i = 0; i < n * n; i++) L[i] = A[i];
    double mn = jacobi_min_eig_n(L, n);
    if (mn < 0)
      for (int k = 0; k < n; k++) A[k + n * k] = A[k + n * k] - mn * 1.0 + 0.01 * 1.0;
    chol_is_spd_n(A, L, n);
  }
  // x = -A^-1 b
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int j = 0; j < i; j++) s -= L[i + n * j] * x[j];
    x[i] = s / L[i + n * i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int j = i + 1; j < n; j++) s -= L[j + n * i] * x[j];
    x[i] = s / L[i + n * i];
  }
  double wolfe = 0;
  for (int i = 0; i < n; i++) { x[i] = -x[i]; }
  for (int i = 0; i < n; i++) wolfe += x[i] * b[i];
  wolfe = -wolfe;
  double dir[18];
  for (int i = 0; i < 18; i++) dir[i] = 0;
  {
    int base = (sp == 0) ? 6 : 0;
    for (int i = 0; i < 3 * tn; i++) dir[base + i] = x[i];
  }
  const double tdir = x[3 * tn];
  double step = 1.0;
  if (t + step * tdir <= 0) step = -0.95 * t / tdir;
  const double e0 = slack_energy_dev(M, cs, ptime, p, t, lam, tlam, a.mu, a.ks, a.kt);
  const double t0 = t;
  t = t0 + step * tdir;
  double pn[18];
  int guard = 0;
  while (guard++ < 400) {
    for (int i = 0; i < 18; i++) pn[i] = p[i] + step * dir[i];
    double e1 = slack_energy_dev(M, cs, ptime, pn, t, lam, tlam, a.mu, a.ks, a.kt);
    if (!(e0 - 1e-4 * wolfe * step < e1)) break;
    step *= 0.8;
    t = t0 + step * tdir;
  }
  for (int i = 0; i < 18; i++) pn[i] = p[i] + step * dir[i];
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 6; r++) {
      size_t s = (size_t)robot * 18 * a.P + (size_t)k * 6 * a.P + 6 * sp + r;
      a.pslack[s] = pn[r * 3 + k];
      a.plambda[s] = lam[r * 3 + k] + a.mu * (cs[r * 3 + k] - pn[r * 3 + k]);
    }
  a.tslack[pb] = t;
  a.tlambda[pb] = tlam + a.mu * (ptime - t);
}

int slack_update(tob_ctx* c, int rb, int re) {
  SlackArgs a;
  a.spline = c->s_spline.p; a.ptime = c->s_ptime.p; a.convert = c->d_convert.p; a.mdyn = c->d_mdyn.p;
  a.pslack = c->s_pslack.p; a.tslack = c->s_tslack.p; a.plambda = c->s_plambda.p; a.tlambda = c->s_tlambda.p;
  a.mu = c->prm.mu; a.ks = c->prm.ks; a.kt = c->prm.kt; a.P = c->prm.piece_num; a.T = c->T; a.robot_begin = rb;
  a.n = (re - rb) * c->prm.piece_num;
  Prof prof(c, K_SLACK);
  k_slack<<<div_up(a.n, 32), 32, 0, c->stream>>>(a);
  TOB_LAUNCH_CHECK(c);
  return 0;
}

}  // namespace tob
