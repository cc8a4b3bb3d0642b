/* port.h -- shared declarations of the plain-C restatement of traj-opt-admm's per-iteration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle_port.so, and only as the checker; the product (traj-opt-admm_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference tree).  Parity of this port
 * is PINNED: tests/test_cpu_checks.py runs it against the tests/golden npz vectors, which were produced by the unmodified reference
 * compiled here (oracle/_ref, tests/golden/make_golden.py).
 *
 * Conventions: matrices are column-major FP64 like Eigen::MatrixXd (spline: T x 3, p_slack / p_lambda: 6P x 3, a
 * sub-segment's control points P: 6 x 3).  No FMA contraction (built with -ffp-contract=off): the reference is built for
 * plain SSE2 (CMakeLists.txt:24).
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H

#define ORDER 5          /* order_num, CCDUtils.h */
#define NCP 6            /* control points per piece */
#define KDOP_AXES 49

typedef struct {
  int piece_num, res, uav_num, n_tr, T;
  double lambda, margin, offset, mu, vel_limit, acc_limit, ks, kt;
  double *basis;   /* n_tr x 36: basis_tr = blossom(a,b) * convert[piece], col-major 6x6 */
  double *weight;  /* n_tr: b - a */
  double *convert; /* piece_num x 36 */
  double mdyn[36];
  double kdop[3 * KDOP_AXES];
  /* cloud */
  int n_pts;
  double *V;       /* n_pts x 3 col-major */
  double wolfe, gnorm;   /* the reference's globals of the same name (CCDUtils.cpp:8-15) */
  /* persistent planes ("optimal_plane": 1): the reference's globals is_seperate / seperate_c / seperate_d (CCDUtils.cpp:34-36,
   * dense N_tr x N_pts, Main/admmPathPlanning3D.cpp:342-351) and is_self_seperate / self_seperate_c / _d (:30-32) */
  int optimal_plane;
  unsigned char *is_sep;      /* n_tr x n_pts */
  double *sep_cd;             /* n_tr x n_pts x 4 */
  unsigned char *is_self_sep; /* n_tr x u x u */
  double *self_sep_cd;        /* n_tr x u x u x 4 */
} port_ctx;

extern port_ctx g_port;

/* port_gjk.c */
void port_gjk_witness(const double (*A)[3], int na, const double (*B)[3], int nb, double *v);

/* port_optplane.c */
int port_optimal_cd_impl(const double (*P)[3], const double *q, double *c, double *d_io);
int port_self_optimal_cd_impl(const double (*P0)[3], const double (*P1)[3], double *c, double *d_io);

/* port_dense.c */
int port_llt(const double *A, double *L, int n);                 /* 0 = Eigen::NumericalIssue */
void port_llt_solve(const double *L, int n, double *x);          /* in place: b -> A^-1 b */
double port_min_eig(const double *A, int n);                     /* smallest eigenvalue (cyclic Jacobi) */

#endif
