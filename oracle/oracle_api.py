"""ctypes front-end of the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (traj-opt-admm_b200/) never does.

  RefOracle  -> oracle/_ref/libtrajopt_ref.so   (the unmodified reference, compiled; see Makefile)
  PortOracle -> oracle/liboracle_port.so        (plain-C restatement, oracle/port/*.c)

Both expose the same method names so parity tests can be parametrised over them.
All matrices are column-major FP64 (numpy order='F'), as Eigen::MatrixXd stores them.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libtrajopt_ref.so")
PORT_SO = os.path.join(HERE, "liboracle_port.so")

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)


def _d(a):
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return a.ctypes.data_as(_dp)


def _u(a):
    assert a.dtype == np.uint32
    return a.ctypes.data_as(_up)


def F(a):
    """column-major float64 copy"""
    return np.array(a, dtype=np.float64, order="F")


class Params:
    """Config File/3D.json values + the constants main() hard-codes (admmPathPlanning3D.cpp:477-478)."""

    def __init__(self, piece_num, res=8, uav_num=1, lam=10.0, margin=0.1, offset=0.1, mu=0.1, vel_limit=2.0,
                 acc_limit=2.0, ks=1e-8, kt=1.0, optimal_plane=0):
        self.piece_num, self.res, self.uav_num = piece_num, res, uav_num
        self.lam, self.margin, self.offset, self.mu = lam, margin, offset, mu
        self.vel_limit, self.acc_limit, self.ks, self.kt = vel_limit, acc_limit, ks, kt
        self.optimal_plane = optimal_plane

    @property
    def n_tr(self):
        return self.piece_num * self.res

    @property
    def T(self):
        return 6 + 3 * (self.piece_num - 1)


class _Base:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.p = None
        self.n_pts = 0

    def _f(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- set-up
    def setup(self, p: Params):
        self.p = p
        self._f("setup")(C.c_int(p.piece_num), C.c_int(p.res), C.c_int(p.uav_num), C.c_double(p.lam),
                         C.c_double(p.margin), C.c_double(p.offset), C.c_double(p.mu), C.c_double(p.vel_limit),
                         C.c_double(p.acc_limit), C.c_double(p.ks), C.c_double(p.kt), C.c_int(p.optimal_plane))

    def tables(self):
        p = self.p
        basis = np.zeros((p.n_tr, 36)); weight = np.zeros(p.n_tr); conv = np.zeros((p.piece_num, 36))
        mdyn = np.zeros(36); kdop = np.zeros(3 * 49)
        self._f("get_tables")(_d(basis), _d(weight), _d(conv), _d(mdyn), _d(kdop))
        return dict(basis=basis, weight=weight, convert=conv, mdyn=mdyn, kdop=kdop)

    def init_pointcloud(self, V):
        V = F(V)
        self.n_pts = V.shape[0]
        self._V = V
        self._f("init_pointcloud")(_d(V), C.c_int(V.shape[0]))

    # ---- broadphase
    def _csr_call(self, fn, args, cap):
        p = self.p
        off = np.zeros(p.n_tr + 1, dtype=np.uint32)
        ids = np.zeros(max(cap, 1), dtype=np.uint32)
        n = fn(*args, _u(off), _u(ids), C.c_long(cap))
        if n > cap:
            return self._csr_call(fn, args, int(n))
        return off, ids[:n]

    def dcd_collision(self, spline, d, cap=1 << 20):
        return self._csr_call(self._f("dcd_collision", C.c_long), (_d(F(spline)), C.c_double(d)), cap)

    def ccd_collision(self, spline, direction, d, cap=1 << 20):
        return self._csr_call(self._f("ccd_collision", C.c_long), (_d(F(spline)), _d(F(direction)), C.c_double(d)), cap)

    def self_dcd(self, P, d):
        P = np.ascontiguousarray(P, dtype=np.float64)  # (u, 18): each 6x3 col-major
        u = P.shape[0]
        pairs = np.zeros(u * u + 2, dtype=np.uint32)
        n = self._f("self_dcd", C.c_long)(_d(P), C.c_int(u), C.c_double(d), _u(pairs), C.c_long(u * u // 2 + 1))
        return pairs[:2 * n].reshape(-1, 2)

    def self_ccd(self, P, D, d):
        P = np.ascontiguousarray(P, dtype=np.float64); D = np.ascontiguousarray(D, dtype=np.float64)
        u = P.shape[0]
        pairs = np.zeros(u * u + 2, dtype=np.uint32)
        n = self._f("self_ccd", C.c_long)(_d(P), _d(D), C.c_int(u), C.c_double(d), _u(pairs), C.c_long(u * u // 2 + 1))
        return pairs[:2 * n].reshape(-1, 2)

    # ---- primitives
    def segment_points(self, spline, tr_id):
        P = np.zeros((6, 3), order="F")
        self._f("segment_points")(_d(F(spline)), C.c_int(tr_id), _d(P))
        return P

    def gjk(self, A, B):
        A = F(np.atleast_2d(A)); B = F(np.atleast_2d(B)); v = np.zeros(3)
        self._f("gjk")(_d(A), C.c_int(A.shape[0]), _d(B), C.c_int(B.shape[0]), _d(v))
        return v

    def kdop_dcd(self, P, q, d):
        return bool(self._f("kdop_dcd", C.c_int)(_d(F(P)), _d(F(q)), C.c_double(d)))

    def self_kdop_dcd(self, P0, P1, d):
        return bool(self._f("self_kdop_dcd", C.c_int)(_d(F(P0)), _d(F(P1)), C.c_double(d)))

    def kdop_ccd(self, P, D, q, d, t0, t1):
        return bool(self._f("kdop_ccd", C.c_int)(_d(F(P)), _d(F(D)), _d(F(q)), C.c_double(d), C.c_double(t0), C.c_double(t1)))

    def gjk_ccd(self, P, D, q, d, t0, t1):
        return bool(self._f("gjk_ccd", C.c_int)(_d(F(P)), _d(F(D)), _d(F(q)), C.c_double(d), C.c_double(t0), C.c_double(t1)))

    def self_kdop_ccd(self, P0, D0, P1, D1, d, t0, t1, s0, s1):
        return bool(self._f("self_kdop_ccd", C.c_int)(_d(F(P0)), _d(F(D0)), _d(F(P1)), _d(F(D1)), C.c_double(d),
                                                       C.c_double(t0), C.c_double(t1), C.c_double(s0), C.c_double(s1)))

    def self_gjk_ccd(self, P0, D0, P1, D1, d, t0, t1, s0, s1):
        return bool(self._f("self_gjk_ccd", C.c_int)(_d(F(P0)), _d(F(D0)), _d(F(P1)), _d(F(D1)), C.c_double(d),
                                                      C.c_double(t0), C.c_double(t1), C.c_double(s0), C.c_double(s1)))

    def opengjk(self, P, q, dist):
        c = np.zeros(3); d = C.c_double(0)
        ok = self._f("opengjk", C.c_int)(_d(F(P)), _d(F(q)), C.c_double(dist), _d(c), C.byref(d))
        return bool(ok), c, d.value

    def selfgjk(self, P0, P1, dist):
        c = np.zeros(3); d = C.c_double(0)
        ok = self._f("selfgjk", C.c_int)(_d(F(P0)), _d(F(P1)), C.c_double(dist), _d(c), C.byref(d))
        return bool(ok), c, d.value

    def optimal_d(self, P0, P1, c, d):
        dd = C.c_double(d)
        self._f("optimal_d")(_d(F(P0)), _d(F(P1)), _d(F(c)), C.byref(dd))
        return dd.value

    # ---- persistent-plane mode (Params.optimal_plane = 1)
    def optimal_cd(self, P, q, c, d):
        """Optimal_plane::optimal_cd: refined (c, d) of one (sub-segment, point) plane"""
        cc = np.array(c, dtype=np.float64); dd = C.c_double(d)
        self._f("optimal_cd")(_d(F(P)), _d(np.ascontiguousarray(q, dtype=np.float64)), _d(cc), C.byref(dd))
        return cc, dd.value

    def self_optimal_cd(self, P0, P1, c, d):
        """Optimal_plane::self_optimal_cd: refined (c, d) of one inter-robot plane"""
        cc = np.array(c, dtype=np.float64); dd = C.c_double(d)
        self._f("self_optimal_cd")(_d(F(P0)), _d(F(P1)), _d(cc), C.byref(dd))
        return cc, dd.value

    def reset_persistent_planes(self):
        """empties is_seperate / is_self_seperate (the state init_variable allocates); call after setup + init_pointcloud"""
        self._f("reset_persistent_planes")()

    def live_planes(self):
        """(tr, point id, c, d) of the live obstacle planes in (tr, id) order"""
        f = self._f("live_planes", C.c_long)
        n = f(None, None, None, None, C.c_long(0))
        tr = np.zeros(max(n, 1), dtype=np.uint32); ids = np.zeros(max(n, 1), dtype=np.uint32)
        c = np.zeros((max(n, 1), 3)); d = np.zeros(max(n, 1))
        f(_u(tr), _u(ids), _d(c), _d(d), C.c_long(n))
        return tr[:n], ids[:n], c[:n], d[:n]

    # ---- planes
    def separate_plane(self, spline, cap=1 << 20):
        p = self.p
        off = np.zeros(p.n_tr + 1, dtype=np.uint32)
        c = np.zeros((max(cap, 1), 3)); d = np.zeros(max(cap, 1))
        n = self._f("separate_plane", C.c_long)(_d(F(spline)), _u(off), _d(c), _d(d), C.c_long(cap))
        if n > cap:
            return self.separate_plane(spline, int(n))
        return off, c[:n].copy(), d[:n].copy()

    def separate_self(self, splines, cap=1 << 18):
        p = self.p
        u = len(splines)
        S = np.concatenate([F(s).ravel(order="F") for s in splines])
        off = np.zeros(u * p.n_tr + 1, dtype=np.uint32)
        c = np.zeros((cap, 3)); d = np.zeros(cap)
        n = self._f("separate_self", C.c_long)(_d(S), C.c_int(u), _u(off), _d(c), _d(d), C.c_long(cap))
        assert n <= cap
        return off, c[:n].copy(), d[:n].copy()

    # ---- energies / gradients
    def _state_args(self, st):
        return (_d(F(st["spline"])), C.c_double(st["piece_time"]), _d(F(st["p_slack"])), _d(F(st["t_slack"])),
                _d(F(st["p_lambda"])), _d(F(st["t_lambda"])))

    @staticmethod
    def _plane_args(planes):
        off, c, d = planes
        return (_u(np.ascontiguousarray(off, dtype=np.uint32)), _d(np.ascontiguousarray(c, dtype=np.float64)),
                _d(np.ascontiguousarray(d, dtype=np.float64)))

    def plane_barrier_energy(self, spline, planes):
        return self._f("plane_barrier_energy", C.c_double)(_d(F(spline)), *self._plane_args(planes))

    def bound_energy(self, spline, piece_time):
        return self._f("bound_energy", C.c_double)(_d(F(spline)), C.c_double(piece_time))

    def spline_energy(self, st, planes):
        return self._f("spline_energy", C.c_double)(*self._state_args(st), *self._plane_args(planes))

    def local_spline_gradient(self, st, planes, sp_id):
        g = np.zeros(19); h = np.zeros((19, 19), order="F")
        self._f("local_spline_gradient")(*self._state_args(st), *self._plane_args(planes), C.c_int(sp_id), _d(g), _d(h))
        return g, h

    def global_spline_gradient(self, st, planes):
        n = 3 * self.p.T + 1
        g = np.zeros(n); h = np.zeros((n, n), order="F")
        self._f("global_spline_gradient")(*self._state_args(st), *self._plane_args(planes), _d(g), _d(h))
        return g, h

    def descent_direction(self, st, planes, multi=False):
        T = self.p.T
        direction = np.zeros((T, 3), order="F")
        td = C.c_double(0); w = C.c_double(0); gn = C.c_double(0)
        name = "descent_direction_multi" if multi else "descent_direction"
        self._f(name)(*self._state_args(st), *self._plane_args(planes), _d(direction), C.byref(td), C.byref(w), C.byref(gn))
        return direction, td.value, w.value, gn.value

    # ---- function-level rows behind the remaining C-ABI entries (reference oracle / host drop-in only, not in the C port)
    def local_plane_barrier_gradient(self, spline, tr_id, c, d):
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1, 3); d = np.ascontiguousarray(d, dtype=np.float64)
        g = np.zeros(18); h = np.zeros((18, 18), order="F")
        self._f("local_plane_barrier_gradient")(_d(F(spline)), C.c_int(tr_id), _d(c), _d(d), C.c_int(len(d)), _d(g), _d(h))
        return g, h

    def local_bound_gradient(self, spline, tr_id, piece_time):
        g = np.zeros(18); h = np.zeros((18, 18), order="F"); pg = np.zeros(18); gt = C.c_double(0); ht = C.c_double(0)
        self._f("local_bound_gradient")(_d(F(spline)), C.c_int(tr_id), C.c_double(piece_time), _d(g), _d(h), C.byref(gt), C.byref(ht), _d(pg))
        return g, h, gt.value, ht.value, pg

    def slack_energy(self, c_spline, piece_time, p_part, t_part, p_lambda, t_lambda):
        return self._f("slack_energy", C.c_double)(_d(F(c_spline)), C.c_double(piece_time), _d(F(p_part)), C.c_double(t_part),
                                                   _d(F(p_lambda)), C.c_double(t_lambda))

    def slack_gradient(self, c_spline, piece_time, p_part, t_part, p_lambda, t_lambda):
        g = np.zeros(19); h = np.zeros((19, 19), order="F")
        self._f("slack_gradient")(_d(F(c_spline)), C.c_double(piece_time), _d(F(p_part)), C.c_double(t_part), _d(F(p_lambda)),
                                  C.c_double(t_lambda), _d(g), _d(h))
        return g, h

    def dynamic_energy(self, p_part, t_part):
        return self._f("dynamic_energy", C.c_double)(_d(F(p_part)), C.c_double(t_part))

    def dynamic_gradient(self, p_part, t_part):
        g = np.zeros(18); h = np.zeros((18, 18), order="F"); pg = np.zeros(18); gt = C.c_double(0); ht = C.c_double(0)
        self._f("dynamic_gradient")(_d(F(p_part)), C.c_double(t_part), _d(g), _d(h), C.byref(gt), C.byref(ht), _d(pg))
        return g, h, gt.value, ht.value, pg

    def line_search(self, st, direction, t_direction, wolfe, planes, step=None):
        """Optimization3D_admm::spline_line_search (step None: bound = Step::position_step) or the multi-UAV variant with the
        caller's bound `step`; returns (spline, piece_time[, accepted step])"""
        sp = F(st["spline"]); pt = C.c_double(st["piece_time"])
        args = (_d(sp), C.byref(pt), _d(F(direction)), C.c_double(t_direction), C.c_double(wolfe), _d(F(st["p_slack"])),
                _d(F(st["t_slack"])), _d(F(st["p_lambda"])), _d(F(st["t_lambda"])), *self._plane_args(planes))
        if step is None:
            self._f("line_search")(*args)
            return sp, pt.value
        s_io = C.c_double(step)
        self._f("line_search_multi")(*args, C.byref(s_io))
        return sp, pt.value, s_io.value

    def edge_collision(self, edge, d, cap=1 << 16):
        ids = np.zeros(cap, dtype=np.uint32)
        n = self._f("edge_collision", C.c_long)(_d(F(edge)), C.c_double(d), _u(ids), C.c_long(cap))
        if n > cap:
            return self.edge_collision(edge, d, int(n))
        return ids[:n]

    def read_obj(self, path, cap=1 << 16):
        """Mesh::readOBJ of the reference (vertex block of an OBJ file, with its quirks)"""
        V = np.zeros((cap, 3), order="F")
        n = self._f("read_obj", C.c_long)(str(path).encode(), _d(V), C.c_long(cap))
        if n > cap:
            return self.read_obj(path, int(n))
        return V[:n].copy(order="F")

    def gjk_dcd(self, A, B, d):
        A = F(np.atleast_2d(A)); B = F(np.atleast_2d(B))
        return bool(self._f("gjk_dcd", C.c_int)(_d(A), C.c_int(A.shape[0]), _d(B), C.c_int(B.shape[0]), C.c_double(d)))

    # ---- steps
    def position_step(self, spline, direction):
        return self._f("position_step", C.c_double)(_d(F(spline)), _d(F(direction)))

    def self_step(self, splines, directions):
        u = len(splines)
        S = np.concatenate([F(s).ravel(order="F") for s in splines])
        D = np.concatenate([F(s).ravel(order="F") for s in directions])
        steps = np.zeros(u)
        self._f("self_step")(_d(S), _d(D), C.c_int(u), _d(steps))
        return steps

    def couple_self_step(self, splines, directions):
        u = len(splines)
        S = np.concatenate([F(s).ravel(order="F") for s in splines])
        D = np.concatenate([F(s).ravel(order="F") for s in directions])
        return self._f("couple_self_step", C.c_double)(_d(S), _d(D), C.c_int(u))

    # ---- slack + whole iteration
    def update_slack_lambda(self, st):
        ps = F(st["p_slack"]); ts = F(st["t_slack"]); pl = F(st["p_lambda"]); tl = F(st["t_lambda"])
        self._f("update_slack_lambda")(_d(F(st["spline"])), C.c_double(st["piece_time"]), _d(ps), _d(ts), _d(pl), _d(tl))
        out = dict(st); out.update(p_slack=ps, t_slack=ts, p_lambda=pl, t_lambda=tl)
        return out

    def optimization(self, st):
        sp = F(st["spline"]); ps = F(st["p_slack"]); ts = F(st["t_slack"]); pl = F(st["p_lambda"]); tl = F(st["t_lambda"])
        pt = C.c_double(st["piece_time"]); gn = C.c_double(0)
        self._f("optimization")(_d(sp), C.byref(pt), _d(ps), _d(ts), _d(pl), _d(tl), C.byref(gn))
        return dict(spline=sp, piece_time=pt.value, p_slack=ps, t_slack=ts, p_lambda=pl, t_lambda=tl, gnorm=gn.value)

    def optimization_multi(self, sts, coupled=False):
        u = len(sts)
        cat = lambda k: np.concatenate([F(s[k]).ravel(order="F") for s in sts])
        S, PS, TS, PL, TL = cat("spline"), cat("p_slack"), cat("t_slack"), cat("p_lambda"), cat("t_lambda")
        PT = np.array([s["piece_time"] for s in sts], dtype=np.float64)
        gn = C.c_double(0)
        self._f("optimization_multi")(C.c_int(int(coupled)), C.c_int(u), _d(S), _d(PT), _d(PS), _d(TS), _d(PL), _d(TL), C.byref(gn))
        T, P = self.p.T, self.p.piece_num
        out = []
        for i in range(u):
            out.append(dict(spline=S[3 * T * i:3 * T * (i + 1)].reshape((T, 3), order="F").copy(order="F"),
                            piece_time=float(PT[0 if coupled else i]),
                            p_slack=PS[18 * P * i:18 * P * (i + 1)].reshape((6 * P, 3), order="F").copy(order="F"),
                            t_slack=TS[P * i:P * (i + 1)].copy(),
                            p_lambda=PL[18 * P * i:18 * P * (i + 1)].reshape((6 * P, 3), order="F").copy(order="F"),
                            t_lambda=TL[P * i:P * (i + 1)].copy(), gnorm=gn.value))
        return out


class RefOracle(_Base):
    prefix = "ref_"
    kind = "reference"

    def __init__(self):
        super().__init__(REF_SO)


class PortOracle(_Base):
    prefix = "port_"
    kind = "port"

    def __init__(self):
        super().__init__(PORT_SO)


def available():
    out = []
    if os.path.exists(REF_SO):
        out.append("ref")
    if os.path.exists(PORT_SO):
        out.append("port")
    return out


def get(kind=None):
    """best available oracle: the compiled reference if present, else the C port"""
    if kind in (None, "ref") and os.path.exists(REF_SO):
        return RefOracle()
    if kind in (None, "port") and os.path.exists(PORT_SO):
        return PortOracle()
    raise FileNotFoundError("no oracle library built (run `make -C oracle`)")
