"""Direct GPU tests of the function-level C-ABI entry points that the iteration tests only cover transitively:

  tob_self_broadphase   BVH::SelfDCDCollision / SelfCCDCollision   (BVH.cpp:252-329)      SURVEY 8a B4, B5
  tob_box_query         BVH::EdgeCollision                         (BVH.cpp:95-133)        8f-3
  tob_gjk_batch (2,1)   CCD::GJKDCD                                (CCD.h:17-114)          8f-3
  tob_edge_validity     the OMPL motion validator's obstacle loop  (OMPL.cpp:36-96)        8f-3, batched on the LBVH
  tob_row_blocks        Gradient_admm::local_plane_barrier_gradient / local_bound_gradient (Gradient_admm.h:331-572)  G1, G2
  tob_slack_terms       Energy_admm::slack_energy / dynamic_energy, Gradient_admm::slack_gradient / dynamic_gradient   E4, G5
  tob_line_search       Optimization3D_admm::spline_line_search :505-557 and the multi-UAV variant :754-811            L4
  tob_descent_direction(dense_shift=1)   Optimization3D_multi::spline_descent_direction :659-752                      L2

Every call is made twice through the SAME caller source (oracle/ref_shim.cpp): once compiled against the reference headers
(oracle/_ref, CPU) and once against the shadow headers of traj-opt-admm_b200/host (C++ drop-in -> C ABI -> GPU).
"""
import os

import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSHIM = os.path.join(ROOT, "traj-opt-admm_b200", "host", "build", "libtrajopt_hostshim.so")
P = 8


class HostDropIn(oa._Base):
    prefix = "ref_"
    kind = "b200-host"

    def __init__(self):
        super().__init__(HOSTSHIM)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def pair(oracle_ref):
    if not os.path.exists(HOSTSHIM):
        pytest.skip("host drop-in not built (needs /root/reference at build time)")
    sc = scenes.bridge(n_pts=20000, seed=31)
    ref, dev = oracle_ref, HostDropIn()
    for o in (ref, dev):
        o.setup(oa.Params(P, ks=sc["ks"]))
        o.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    for _ in range(3):
        st = ref.optimization(st)
    return dict(ref=ref, dev=dev, st=st, sc=sc)


def pair_set(p):
    return sorted((int(a), int(b)) for a, b in p)


def test_self_broadphase_dcd_and_ccd(pair):
    """B4 / B5: per time slot, all pairs of robots whose (swept) boxes are within d; first < second"""
    ref, dev = pair["ref"], pair["dev"]
    rng = np.random.default_rng(17)
    n_pairs = 0
    for u in (2, 8, 33, 64):
        ctr = rng.uniform(-1.5, 1.5, size=(u, 1, 3)) * np.array([1.0, 1.0, 0.2])
        Pm = ctr + rng.normal(size=(u, 6, 3)) * 0.15
        Dm = rng.normal(size=(u, 6, 3)) * 0.3
        Pf = np.ascontiguousarray(np.transpose(Pm, (0, 2, 1)).reshape(u, 18))     # each 6x3 column-major
        Df = np.ascontiguousarray(np.transpose(Dm, (0, 2, 1)).reshape(u, 18))
        for d in (0.1, 0.3):
            a, b = ref.self_dcd(Pf, d), dev.self_dcd(Pf, d)
            assert all(x < y for x, y in b)
            assert pair_set(a) == pair_set(b)
            n_pairs += len(a)
            a, b = ref.self_ccd(Pf, Df, d), dev.self_ccd(Pf, Df, d)
            assert pair_set(a) == pair_set(b)
            n_pairs += len(a)
    assert n_pairs > 100


def test_edge_collision_and_gjk_dcd(pair):
    """8f-3: BVH::EdgeCollision candidate set + CCD::GJKDCD decisions of the front end's path validation"""
    ref, dev, sc = pair["ref"], pair["dev"], pair["sc"]
    rng = np.random.default_rng(5)
    V = sc["V"]
    n_hit = n_col = 0
    for _ in range(24):
        a = np.array([rng.uniform(-6, 6), rng.uniform(-0.6, 0.6), rng.uniform(-0.3, 0.5)])
        b = a + rng.normal(size=3) * np.array([0.6, 0.2, 0.2])
        edge = np.asfortranarray(np.stack([a, b]))
        d = 0.15
        ra, rb = ref.edge_collision(edge, d), dev.edge_collision(edge, d)
        assert np.array_equal(np.sort(ra), rb)
        n_hit += len(ra)
        for pid in ra[:40]:
            q = V[pid:pid + 1]
            x, y = ref.gjk_dcd(edge, q, d), dev.gjk_dcd(edge, q, d)
            assert x == y
            n_col += x
    assert n_hit > 200 and n_col > 0


def test_edge_validity_batch(pair):
    """the motion validator's obstacle loop (OMPL.cpp:36-96: EdgeCollision, then GJKDCD per candidate, invalid on the first
    collision), batched over many edges in one launch on the LBVH"""
    ref, sc = pair["ref"], pair["sc"]
    s = api.Solver(P, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    rng = np.random.default_rng(9)
    n = 300
    a = np.column_stack([rng.uniform(-6.5, 6.5, n), rng.uniform(-0.8, 0.8, n), rng.uniform(-0.5, 0.8, n)])
    b = a + rng.normal(size=(n, 3)) * np.array([0.8, 0.25, 0.25])
    edges = np.stack([a, b], axis=1)            # (n, 2, 3)
    d = 0.15
    got = s.edge_validity(edges, d)
    exp = np.ones(n, dtype=bool)
    for i in range(n):
        e = np.asfortranarray(edges[i])
        for pid in ref.edge_collision(e, d):
            if ref.gjk_dcd(e, sc["V"][pid:pid + 1], d):
                exp[i] = False
                break
    assert np.array_equal(got, exp)
    assert 0.1 < exp.mean() < 0.95
    s.close()


def test_row_blocks_plane_and_bound(pair):
    """G1 / G2 separately, in the 18 piece coordinates"""
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    off, c, d = ref.separate_plane(st["spline"])
    busy = [tr for tr in range(P * 8) if off[tr + 1] - off[tr] > 20][:6]
    assert busy
    for tr in busy + [0, P * 8 - 1]:
        cc, dd = c[off[tr]:off[tr + 1]], d[off[tr]:off[tr + 1]]
        g0, h0 = ref.local_plane_barrier_gradient(st["spline"], tr, cc, dd)
        g1, h1 = dev.local_plane_barrier_gradient(st["spline"], tr, cc, dd)
        if np.abs(h0).max() > 0:
            assert rel(g1, g0) < 1e-9 and rel(h1, h0) < 1e-9, tr
        else:
            assert np.abs(h1).max() == 0 and np.abs(g1).max() == 0
    n_active = 0
    # short piece times switch the velocity / acceleration barriers on (inf below the shortest feasible time: skipped)
    # (the barrier is active in the narrow band vel_limit - margin < v < vel_limit: a fine grid of times over all rows)
    for t in [st["piece_time"]] + list(np.linspace(1.7, 2.2, 26)):
        for tr in range(0, P * 8, 3):
            a = ref.local_bound_gradient(st["spline"], tr, t)
            b = dev.local_bound_gradient(st["spline"], tr, t)
            if not np.all(np.isfinite(a[0])):
                continue
            sc_ = max(np.abs(a[1]).max(), 1e-300)
            n_active += np.abs(a[1]).max() > 0
            assert np.max(np.abs(b[0] - a[0])) <= 1e-9 * max(np.abs(a[0]).max(), 1e-300) + 0
            assert np.max(np.abs(b[1] - a[1])) <= 1e-9 * sc_
            assert abs(b[2] - a[2]) <= 1e-9 * max(abs(a[2]), 1e-300) and abs(b[3] - a[3]) <= 1e-9 * max(abs(a[3]), 1e-300)
            assert np.max(np.abs(b[4] - a[4])) <= 1e-9 * max(np.abs(a[4]).max(), 1e-300)
    assert n_active > 0


def test_slack_terms(pair):
    """E4 / G5: one piece of the slack problem, with and without the consensus terms"""
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    rng = np.random.default_rng(2)
    for sp in (0, 3, P - 1):
        cs = np.asfortranarray(st["spline"][3 * sp:3 * sp + 6] + rng.normal(size=(6, 3)) * 0.01)
        pp = np.asfortranarray(st["p_slack"][6 * sp:6 * sp + 6])
        pl = np.asfortranarray(st["p_lambda"][6 * sp:6 * sp + 6] + rng.normal(size=(6, 3)) * 0.01)
        t, tp, tl = st["piece_time"], float(st["t_slack"][sp]), float(st["t_lambda"][sp]) + 0.01
        e0, e1 = ref.slack_energy(cs, t, pp, tp, pl, tl), dev.slack_energy(cs, t, pp, tp, pl, tl)
        assert abs(e1 - e0) <= 1e-9 * abs(e0)
        (g0, h0), (g1, h1) = ref.slack_gradient(cs, t, pp, tp, pl, tl), dev.slack_gradient(cs, t, pp, tp, pl, tl)
        assert rel(g1, g0) < 1e-9 and rel(h1, h0) < 1e-9
        e0, e1 = ref.dynamic_energy(pp, tp), dev.dynamic_energy(pp, tp)
        assert abs(e1 - e0) <= 1e-9 * abs(e0)
        a, b = ref.dynamic_gradient(pp, tp), dev.dynamic_gradient(pp, tp)
        for x, y in zip(a, b):
            assert rel(y, x) < 1e-9 or np.max(np.abs(np.asarray(x))) == 0


def test_line_search_single_and_multi(pair):
    """L4: CCD-bounded Armijo search of the single-UAV path, and the multi-UAV variant with a caller-provided bound"""
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    planes = ref.separate_plane(st["spline"])
    direction, td, w, gn = ref.descent_direction(st, planes)
    sp0, pt0 = ref.line_search(st, direction, td, w, planes)
    sp1, pt1 = dev.line_search(st, direction, td, w, planes)
    assert np.max(np.abs(sp0 - st["spline"])) > 0                      # a step was taken
    assert np.max(np.abs(sp1 - sp0)) <= 1e-12 * max(1.0, np.abs(sp0).max()) and abs(pt1 - pt0) <= 1e-12 * abs(pt0)
    # a direction that is far too long: the ladder has to back off several rungs (CCD bound and Armijo)
    big = np.asfortranarray(8.0 * direction)
    sp0, pt0 = ref.line_search(st, big, 8.0 * td, 8.0 * w, planes)
    sp1, pt1 = dev.line_search(st, big, 8.0 * td, 8.0 * w, planes)
    assert np.max(np.abs(sp1 - sp0)) <= 1e-12 * max(1.0, np.abs(sp0).max()) and abs(pt1 - pt0) <= 1e-12 * abs(pt0)
    for bound in (1.0, 0.64, 0.8 ** 7):
        a = ref.line_search(st, direction, td, w, planes, step=bound)
        b = dev.line_search(st, direction, td, w, planes, step=bound)
        assert a[2] == b[2] and a[2] <= bound
        assert np.max(np.abs(b[0] - a[0])) <= 1e-12 * max(1.0, np.abs(a[0]).max()) and abs(b[1] - a[1]) <= 1e-12 * abs(a[1])


def test_descent_direction_multi_variant(pair):
    """L2: the multi-UAV direction (dense LLT with the global eigen-shift fall-back, Optimization3D_multi.h:659-752); also at
    piece times where piece blocks are indefinite before the per-piece projection"""
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    planes = ref.separate_plane(st["spline"])
    tested = 0
    for t in (st["piece_time"], 2.6, 2.2):
        st2 = dict(st, piece_time=float(t))
        if not np.isfinite(ref.spline_energy(st2, planes)):
            continue
        d0, td0, w0, gn0 = ref.descent_direction(st2, planes, multi=True)
        d1, td1, w1, gn1 = dev.descent_direction(st2, planes, multi=True)
        assert rel(d1, d0) < 1e-7 and abs(td1 - td0) <= 1e-7 * max(abs(td0), 1e-12)
        assert abs(w1 - w0) <= 1e-8 * abs(w0) and abs(gn1 - gn0) <= 1e-9 * abs(gn0)
        tested += 1
    assert tested >= 2
