"""ctypes binding of libtrajopt_b200.so (C ABI: include/trajopt_b200.h).

Method names mirror the reference's entry points so parity tests read like calls into the reference:
  BVH::InitPointcloud -> init_pointcloud, BVH::DCDCollision -> dcd_collision, BVH::CCDCollision -> ccd_collision,
  Optimization3D_admm::separate_plane -> separate_plane, Energy_admm::spline_energy -> spline_energy,
  Gradient_admm::global_spline_gradient -> global_spline_gradient, Step::position_step -> position_step,
  Optimization3D_admm::optimization -> optimization, ...
There is NO CPU fallback: if the shared library is missing or no GPU is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "libtrajopt_b200.so")

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(C.c_uint8)


class TobParams(C.Structure):
    _fields_ = [("piece_num", C.c_int32), ("res", C.c_int32), ("uav_num", C.c_int32), ("optimal_plane", C.c_int32),
                ("lam", C.c_double), ("margin", C.c_double), ("offset", C.c_double), ("mu", C.c_double),
                ("vel_limit", C.c_double), ("acc_limit", C.c_double), ("ks", C.c_double), ("kt", C.c_double)]


class TobState(C.Structure):
    _fields_ = [("spline", _dp), ("piece_time", _dp), ("p_slack", _dp), ("t_slack", _dp), ("p_lambda", _dp), ("t_lambda", _dp)]


class TobCounters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("dcd_candidates", C.c_uint64), ("planes", C.c_uint64),
                ("ccd_candidates", C.c_uint64), ("energy_plane_evals", C.c_uint64), ("self_pairs", C.c_uint64),
                ("line_search_trials", C.c_uint64), ("barrier_terms", C.c_uint64), ("live_planes", C.c_uint64),
                ("refine_capped", C.c_uint64), ("np_kdop_groups", C.c_uint64), ("np_gjk_iters", C.c_uint64),
                ("ccd_gjk_iters", C.c_uint64), ("ccd_kdop_pass", C.c_uint64), ("np_kdop_exact", C.c_uint64), ("np_band", C.c_uint64), ("ls_rung_hist", C.c_uint64 * 8), ("ls_rungs_skipped", C.c_uint64)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p)


def _d(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def _u(a):
    assert a.dtype == np.uint32
    return a.ctypes.data_as(_up)


def F(a):
    return np.array(a, dtype=np.float64, order="F")


def Fview(a):
    """column-major float64 view when the array already is one (no copy), else a copy"""
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous:
        return a
    return F(a)


def load_library():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libtrajopt_b200.so is not built (run `make -C traj-opt-admm_b200`); there is no CPU fallback")
    return C.CDLL(LIB_PATH)


class _HostState:
    """keeps the numpy buffers a tob_state points to alive"""

    def __init__(self, st, inplace=False):
        # inplace: the caller's own column-major float64 arrays are handed to the library and updated in place (what a C++
        # caller does with its Eigen matrices); otherwise private copies are made and the inputs stay untouched
        G = (lambda a: a if (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous) else F(a)) if inplace else F
        self.spline = G(st["spline"]); self.pt = np.array([st["piece_time"]], dtype=np.float64)
        self.p_slack = G(st["p_slack"]); self.t_slack = G(st["t_slack"])
        self.p_lambda = G(st["p_lambda"]); self.t_lambda = G(st["t_lambda"])
        self.c = TobState(_d(self.spline), _d(self.pt), _d(self.p_slack), _d(self.t_slack), _d(self.p_lambda), _d(self.t_lambda))

    def to_dict(self, **extra):
        d = dict(spline=self.spline, piece_time=float(self.pt[0]), p_slack=self.p_slack, t_slack=self.t_slack,
                 p_lambda=self.p_lambda, t_lambda=self.t_lambda)
        d.update(extra)
        return d


class Solver:
    def __init__(self, piece_num, res=8, uav_num=1, lam=10.0, margin=0.1, offset=0.1, mu=0.1, vel_limit=2.0, acc_limit=2.0,
                 ks=1e-8, kt=1.0, device=0, optimal_plane=0):
        self.lib = load_library()
        self.lib.tob_last_error.restype = C.c_char_p
        self.lib.tob_stream.restype = C.c_void_p
        self.lib.tob_cloud_size.restype = C.c_uint32
        self.ctx = C.c_void_p()
        if self.lib.tob_ctx_create(C.c_int(device), C.byref(self.ctx)):
            raise RuntimeError("tob_ctx_create: " + self.lib.tob_last_error(None).decode())
        self.piece_num, self.res, self.uav_num = piece_num, res, uav_num
        self.n_tr = piece_num * res
        self.T = 6 + 3 * (piece_num - 1)
        self.prm = TobParams(piece_num, res, uav_num, int(optimal_plane), lam, margin, offset, mu, vel_limit, acc_limit, ks, kt)
        self._ck(self.lib.tob_set_params(self.ctx, C.byref(self.prm)))
        self._ck(self.lib.tob_make_tables(self.ctx, None))
        self._cb = None

    def close(self):
        if self.ctx:
            self.lib.tob_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.lib.tob_last_error(self.ctx).decode())

    # ---- set-up
    def tables(self):
        basis = np.zeros((self.n_tr, 36)); weight = np.zeros(self.n_tr); conv = np.zeros((self.piece_num, 36))
        mdyn = np.zeros(36); kdop = np.zeros(147)
        self._ck(self.lib.tob_get_tables(self.ctx, _d(basis), _d(weight), _d(conv), _d(mdyn), _d(kdop)))
        return dict(basis=basis, weight=weight, convert=conv, mdyn=mdyn, kdop=kdop)

    def set_tables(self, t):
        self._ck(self.lib.tob_set_tables(self.ctx, _d(np.ascontiguousarray(t["basis"])), _d(np.ascontiguousarray(t["weight"])),
                                         _d(np.ascontiguousarray(t["convert"])), _d(np.ascontiguousarray(t["mdyn"])),
                                         _d(np.ascontiguousarray(t["kdop"]))))

    def init_pointcloud(self, V):
        V = Fview(V)
        self._ck(self.lib.tob_cloud_upload(self.ctx, _d(V), C.c_uint32(V.shape[0])))

    def build_stats(self):
        """(device ms, points) of the last LBVH build"""
        ms = C.c_double(0); n = C.c_uint64(0)
        self.lib.tob_build_stats(self.ctx, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def init_pointclouds(self, Vs):
        """one cloud per robot slot (batched independent problems, mode 2)"""
        Vs = [Fview(V) for V in Vs]
        ptrs = (_dp * len(Vs))(*[_d(V) for V in Vs])
        ns = (C.c_uint32 * len(Vs))(*[V.shape[0] for V in Vs])
        self._ck(self.lib.tob_cloud_upload_batch(self.ctx, ptrs, ns, C.c_int(len(Vs))))

    def device_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.tob_device_info(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return dict(sm_count=a.value, cc=(b.value, c.value))

    # ---- broadphase
    def _cat(self, xs):
        if isinstance(xs, np.ndarray) and xs.ndim == 2:
            xs = [xs]
        return np.concatenate([F(x).ravel(order="F") for x in xs]), len(xs)

    def _bp(self, fn, args, n_robots, cap=1 << 20):
        off = np.zeros(n_robots * self.n_tr + 1, dtype=np.uint32)
        ids = np.zeros(max(cap, 1), dtype=np.uint32)
        total = C.c_uint64(0)
        self._ck(fn(self.ctx, *args, _u(off), _u(ids), C.c_uint64(cap), C.byref(total)))
        if total.value > cap:
            return self._bp(fn, args, n_robots, int(total.value))
        return off, ids[:total.value]

    def dcd_collision(self, splines, d):
        S, n = self._cat(splines)
        return self._bp(self.lib.tob_broadphase_dcd, (_d(S), C.c_int(n), C.c_double(d)), n)

    def ccd_collision(self, splines, directions, d):
        S, n = self._cat(splines); Dd, _ = self._cat(directions)
        return self._bp(self.lib.tob_broadphase_ccd, (_d(S), _d(Dd), C.c_int(n), C.c_double(d)), n)

    def self_broadphase(self, P, D, d):
        P = np.ascontiguousarray(P, dtype=np.float64); u = P.shape[0]
        Dp = _d(np.ascontiguousarray(D, dtype=np.float64)) if D is not None else None
        pairs = np.zeros(u * u + 2, dtype=np.uint32); total = C.c_uint64(0)
        self._ck(self.lib.tob_self_broadphase(self.ctx, _d(P), Dp, C.c_int(u), C.c_double(d), _u(pairs), C.c_uint64(u * u // 2 + 1),
                                              C.byref(total)))
        return pairs[:2 * total.value].reshape(-1, 2)

    def edge_validity(self, edges, d):
        """edges: (n, 2, 3); True where no cloud point is within d of the segment (the motion validator's obstacle test)"""
        E = np.ascontiguousarray(edges, dtype=np.float64)
        n = E.shape[0]; out = np.zeros(max(n, 1), dtype=np.uint8)
        self._ck(self.lib.tob_edge_validity_batch(self.ctx, _d(E), C.c_int(n), C.c_double(d), out.ctypes.data_as(_bp)))
        return out[:n].astype(bool)

    # ---- batched primitives (inputs: (n, rows, 3) arrays)
    @staticmethod
    def _pack(A):
        A = np.asarray(A, dtype=np.float64)
        return np.ascontiguousarray(np.transpose(A, (0, 2, 1)))  # each block column-major

    def gjk_batch(self, A, B):
        A = np.asarray(A, dtype=np.float64); B = np.asarray(B, dtype=np.float64)
        n = A.shape[0]; v = np.zeros((n, 3))
        self._ck(self.lib.tob_gjk_batch(self.ctx, _d(self._pack(A)), C.c_int(A.shape[1]), _d(self._pack(B)), C.c_int(B.shape[1]),
                                        C.c_int(n), _d(v)))
        return v

    def kdop_dcd_batch(self, P, q, d):
        n = len(P); fl = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.tob_kdop_dcd_batch(self.ctx, _d(self._pack(P)), _d(np.ascontiguousarray(q, dtype=np.float64)), C.c_int(n),
                                             C.c_double(d), fl.ctypes.data_as(_bp)))
        return fl.astype(bool)

    def plane_point_batch(self, P, q, dist):
        n = len(P); ok = np.zeros(n, dtype=np.uint8); c = np.zeros((n, 3)); d = np.zeros(n)
        self._ck(self.lib.tob_plane_point_batch(self.ctx, _d(self._pack(P)), _d(np.ascontiguousarray(q, dtype=np.float64)), C.c_int(n),
                                                C.c_double(dist), ok.ctypes.data_as(_bp), _d(c), _d(d)))
        return ok.astype(bool), c, d

    def plane_hulls_batch(self, P0, P1, dist, refine=True):
        n = len(P0); ok = np.zeros(n, dtype=np.uint8); c = np.zeros((n, 3)); d = np.zeros(n)
        self._ck(self.lib.tob_plane_hulls_batch(self.ctx, _d(self._pack(P0)), _d(self._pack(P1)), C.c_int(n), C.c_double(dist),
                                                C.c_int(int(refine)), ok.ctypes.data_as(_bp), _d(c), _d(d)))
        return ok.astype(bool), c, d

    def optimal_cd_batch(self, P, q, c, d):
        """Optimal_plane::optimal_cd on n (P 6x3, q) pairs starting from planes (c, d); returns (c, d, capped)"""
        n = len(P); c = np.ascontiguousarray(c, dtype=np.float64).copy(); d = np.ascontiguousarray(d, dtype=np.float64).copy()
        q = np.ascontiguousarray(q, dtype=np.float64); cap = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.tob_optimal_cd_batch(self.ctx, _d(self._pack(P)), _d(q), C.c_int(n), _d(c), _d(d), cap.ctypes.data_as(_bp)))
        return c, d, cap.astype(bool)

    def self_optimal_cd_batch(self, P0, P1, c, d):
        """Optimal_plane::self_optimal_cd on n inter-robot pairs; returns (c, d, capped)"""
        n = len(P0); c = np.ascontiguousarray(c, dtype=np.float64).copy(); d = np.ascontiguousarray(d, dtype=np.float64).copy()
        cap = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.tob_self_optimal_cd_batch(self.ctx, _d(self._pack(P0)), _d(self._pack(P1)), C.c_int(n), _d(c), _d(d),
                                                    cap.ctypes.data_as(_bp)))
        return c, d, cap.astype(bool)

    # ---- persistent planes (optimal_plane=1)
    def planes_reset(self):
        self._ck(self.lib.tob_planes_reset(self.ctx))

    def live_planes(self):
        """(rows, original point ids, c, d) of the live obstacle planes"""
        total = C.c_uint64(0)
        self._ck(self.lib.tob_live_planes(self.ctx, None, None, None, None, C.c_uint64(0), C.byref(total)))
        n = int(total.value)
        rows = np.zeros(max(n, 1), dtype=np.uint32); ids = np.zeros(max(n, 1), dtype=np.uint32)
        c = np.zeros((max(n, 1), 3)); d = np.zeros(max(n, 1))
        self._ck(self.lib.tob_live_planes(self.ctx, _u(rows), _u(ids), _d(c), _d(d), C.c_uint64(n), C.byref(total)))
        return rows[:n], ids[:n], c[:n], d[:n]

    # ---- planes
    def separate_planes(self, splines, with_self=False, cap=1 << 20):
        S, n = self._cat(splines)
        off = np.zeros(n * self.n_tr + 1, dtype=np.uint32)
        c = np.zeros((max(cap, 1), 3)); d = np.zeros(max(cap, 1)); total = C.c_uint64(0)
        self._ck(self.lib.tob_separate_planes(self.ctx, _d(S), C.c_int(n), C.c_int(int(with_self)), _u(off), _d(c), _d(d),
                                              C.c_uint64(cap), C.byref(total)))
        if total.value > cap:
            return self.separate_planes(splines, with_self, int(total.value))
        return off, c[:total.value].copy(), d[:total.value].copy()

    def separate_plane(self, spline):
        return self.separate_planes([spline], False)

    def set_planes(self, planes, n_robots=1):
        off, c, d = planes
        self._ck(self.lib.tob_set_planes(self.ctx, C.c_int(n_robots), _u(np.ascontiguousarray(off, dtype=np.uint32)),
                                         _d(np.ascontiguousarray(c, dtype=np.float64)), _d(np.ascontiguousarray(d, dtype=np.float64))))

    # ---- energies / gradients (against the resident plane set)
    def plane_barrier_energy(self, spline, robot=0):
        e = C.c_double(0)
        self._ck(self.lib.tob_plane_barrier_energy(self.ctx, C.c_int(robot), _d(F(spline)), C.byref(e)))
        return e.value

    def bound_energy(self, spline, piece_time):
        e = C.c_double(0)
        self._ck(self.lib.tob_bound_energy(self.ctx, _d(F(spline)), C.c_double(piece_time), C.byref(e)))
        return e.value

    def spline_energy(self, st, robot=0):
        hs = _HostState(st); e = C.c_double(0)
        self._ck(self.lib.tob_spline_energy(self.ctx, C.c_int(robot), C.byref(hs.c), C.byref(e)))
        return e.value

    def piece_blocks(self, st, robot=0, project_psd=False):
        hs = _HostState(st); P = self.piece_num
        g = np.zeros((P, 19)); h = np.zeros((P, 361))
        self._ck(self.lib.tob_piece_blocks(self.ctx, C.c_int(robot), C.byref(hs.c), C.c_int(int(project_psd)), _d(g), _d(h)))
        return g, h.reshape(P, 19, 19).transpose(0, 2, 1).copy()

    def global_spline_gradient(self, st, robot=0):
        hs = _HostState(st); n = 3 * self.T + 1
        g = np.zeros(n); h = np.zeros((n, n), order="F")
        self._ck(self.lib.tob_global_gradient(self.ctx, C.c_int(robot), C.byref(hs.c), _d(g), _d(h)))
        return g, h

    def descent_direction(self, st, robot=0, multi=False):
        hs = _HostState(st)
        direction = np.zeros((self.T, 3), order="F"); td = C.c_double(0); w = C.c_double(0); gn = C.c_double(0)
        self._ck(self.lib.tob_descent_direction(self.ctx, C.c_int(robot), C.byref(hs.c), C.c_int(int(multi)), _d(direction),
                                                C.byref(td), C.byref(w), C.byref(gn)))
        return direction, td.value, w.value, gn.value

    # ---- steps
    def position_step(self, spline, direction):
        s = C.c_double(0)
        self._ck(self.lib.tob_position_step(self.ctx, _d(F(spline)), _d(F(direction)), C.byref(s)))
        return s.value

    def self_step(self, splines, directions, coupled=False):
        S, n = self._cat(splines); Dd, _ = self._cat(directions)
        steps = np.zeros(n)
        self._ck(self.lib.tob_self_step(self.ctx, _d(S), _d(Dd), C.c_int(n), C.c_int(int(coupled)), _d(steps)))
        return steps[0] if coupled else steps

    # ---- slack + iteration
    def update_slack_lambda(self, st):
        hs = _HostState(st)
        self._ck(self.lib.tob_update_slack_lambda(self.ctx, C.byref(hs.c)))
        return hs.to_dict()

    def _states(self, sts, inplace=False):
        hss = [_HostState(s, inplace) for s in sts]
        arr = (TobState * len(hss))(*[h.c for h in hss])
        return hss, arr

    def optimization(self, st, mode=0, coupled=False, inplace=False):
        """host in / host out, the shape of Optimization3D_admm::optimization / _multi::optimization_decouple;
        coupled=True (mode 1): Optimization3D_multi::optimization, one shared piece time"""
        if coupled:
            mode = 1
        single = isinstance(st, dict)
        sts = [st] if single else st
        hss, arr = self._states(sts, inplace)
        gn = C.c_double(0)
        self._ck(self.lib.tob_optimization(self.ctx, arr, C.c_int(len(sts)), C.c_int(mode), C.byref(gn)))
        out = [h.to_dict(gnorm=gn.value) for h in hss]
        return out[0] if single else out

    def bind_states(self, sts):
        """keep the tob_state array over the caller's own column-major float64 buffers (what a C++ caller holds anyway):
        optimization_bound() then costs one C call, no per-call marshalling"""
        return self._states(sts, inplace=True)

    def optimization_bound(self, bound, mode=0):
        hss, arr = bound
        gn = C.c_double(0)
        self._ck(self.lib.tob_optimization(self.ctx, arr, C.c_int(len(hss)), C.c_int(mode), C.byref(gn)))
        return gn.value

    def states_upload(self, sts):
        hss, arr = self._states(sts)
        self._ck(self.lib.tob_states_upload(self.ctx, arr, C.c_int(len(sts))))

    def states_download(self, sts_like):
        hss, arr = self._states(sts_like)
        self._ck(self.lib.tob_states_download(self.ctx, arr, C.c_int(len(hss))))
        return [h.to_dict() for h in hss]

    def iterate(self, iters=1, mode=0):
        gn = C.c_double(0)
        self._ck(self.lib.tob_admm_iterate(self.ctx, C.c_int(iters), C.c_int(mode), C.byref(gn)))
        return gn.value

    def counters(self):
        c = TobCounters()
        self.lib.tob_get_counters(self.ctx, C.byref(c))
        return {k: (list(getattr(c, k)) if k == "ls_rung_hist" else getattr(c, k)) for k, _ in TobCounters._fields_}

    def reset_counters(self):
        self.lib.tob_reset_counters(self.ctx)

    def profile_enable(self, on=True):
        self._ck(self.lib.tob_profile_enable(self.ctx, C.c_int(int(on))))

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated since profile_enable(True)"""
        out = {}
        kid = 0
        while True:
            ms = C.c_double(0); n = C.c_uint64(0); name = C.c_char_p()
            if self.lib.tob_profile_read(self.ctx, C.c_int(kid), C.byref(ms), C.byref(n), C.byref(name)):
                break
            out[name.value.decode()] = (ms.value, n.value)
            kid += 1
        return out

    def stream(self):
        return self.lib.tob_stream(self.ctx)

    def fp64_peak_tflops(self):
        t = C.c_double(0)
        self._ck(self.lib.tob_fp64_peak(self.ctx, C.byref(t)))
        return t.value

    # ---- multi-GPU: robots sharded over the ranks of a NCCL communicator (see include/trajopt_b200.h)
    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        if self.lib.tob_nccl_unique_id(buf):
            raise RuntimeError("tob_nccl_unique_id: " + self.lib.tob_last_error(None).decode())
        return bytes(buf)

    def nccl_init_rank(self, uid, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.tob_nccl_init_rank(self.ctx, buf, C.c_int(rank), C.c_int(world)))
        return self.shard_range()

    def nccl_detach(self):
        self._ck(self.lib.tob_nccl_detach(self.ctx))

    def shard_range(self):
        a, b, r, w = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.tob_shard_range(self.ctx, C.byref(a), C.byref(b), C.byref(r), C.byref(w))
        return a.value, b.value

    def set_shard(self, first, count, allgather, allreduce):
        self._cb = (ALLGATHER_FN(allgather), ALLREDUCE_FN(allreduce))
        self._ck(self.lib.tob_set_shard(self.ctx, C.c_int(first), C.c_int(count), C.c_int(self.uav_num), self._cb[0], self._cb[1], None))
