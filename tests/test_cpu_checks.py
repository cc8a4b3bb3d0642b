"""CPU-side tests (run with -m "not gpu"): the oracle against the golden vectors generated from the compiled
reference, the host build of the product's math header against the same vectors (bit-exact), and the C ABI surface.
No compute call of the CUDA library is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle_api as oa
from trajopt import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dp = C.POINTER(C.c_double)


def D(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(GOLD, "single.npz"))


@pytest.fixture(scope="module")
def gm():
    return np.load(os.path.join(GOLD, "multi.npz"))


def oracles():
    return [k for k in ("ref", "port") if k in oa.available()]


def state_at(g, i, prefix=""):
    return {k: g["it%d_%s%s" % (i, prefix, k)] for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda")} | \
        {"piece_time": float(g["it%d_%spiece_time" % (i, prefix)])}


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_single_against_golden(gs, kind):
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    exact = kind == "ref"
    o = oa.get(kind)
    P = 4
    o.setup(oa.Params(P, ks=float(gs["ks"])))
    o.init_pointcloud(gs["V"])
    tb = o.tables()
    for k in tb:
        assert np.array_equal(tb[k], gs["tab_" + k]), k
    st = state_at(gs, 2)
    off, ids = o.dcd_collision(st["spline"], 0.2)
    assert np.array_equal(off, gs["dcd_off"])
    got = np.concatenate([np.sort(ids[off[r]:off[r + 1]]) for r in range(len(off) - 1)])
    assert np.array_equal(got, gs["dcd_ids"])
    po, pc, pd = o.separate_plane(st["spline"])
    assert np.array_equal(po, gs["pl_off"])
    planes = (gs["pl_off"], gs["pl_c"], gs["pl_d"])
    tol = 0 if exact else 1e-12
    assert abs(o.spline_energy(st, planes) - float(gs["e_spline"])) <= tol * abs(float(gs["e_spline"]))
    assert abs(o.bound_energy(st["spline"], st["piece_time"]) - float(gs["e_bound"])) <= tol * max(abs(float(gs["e_bound"])), 1e-30)
    g, h = o.global_spline_gradient(st, planes)
    assert np.max(np.abs(g - gs["grad"])) <= (0 if exact else 1e-9) * np.max(np.abs(gs["grad"]))
    assert np.max(np.abs(h - gs["hess"])) <= (0 if exact else 1e-9) * np.max(np.abs(gs["hess"]))
    for dd, sref in zip(gs["step_dirs"], gs["steps"]):
        s = o.position_step(st["spline"], dd)
        assert s <= sref and s == sref
    # iterations
    cur = state_at(gs, 0)
    for it in range(1, 5):
        cur = o.optimization(cur)
        assert np.max(np.abs(cur["spline"] - gs["it%d_spline" % it])) <= (0 if exact else 1e-6)
        assert abs(cur["piece_time"] - float(gs["it%d_piece_time" % it])) <= (0 if exact else 1e-6)


@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_multi_against_golden(gm, kind):
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    exact = kind == "ref"
    o = oa.get(kind)
    U, P = 4, 4
    o.setup(oa.Params(P, uav_num=U, ks=float(gm["ks"])))
    o.init_pointcloud(gm["V"])
    sts = [state_at(gm, 0, "u%d_" % u) for u in range(U)]
    for it in range(1, 4):
        sts = o.optimization_multi(sts, coupled=False)
        for u in range(U):
            assert np.max(np.abs(sts[u]["spline"] - gm["it%d_u%d_spline" % (it, u)])) <= (0 if exact else 1e-6)
    splines = [gm["it2_u%d_spline" % u] for u in range(U)]
    steps = o.self_step(splines, list(gm["self_dirs"]))
    assert np.all(steps <= gm["self_steps"])
    assert np.array_equal(steps, gm["self_steps"])
    assert o.couple_self_step(splines, list(gm["self_dirs"])) == float(gm["couple_step"])


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_coupled_multi_against_golden(kind):
    """Optimization3D_multi::optimization (Optimization3D_multi.h:120-174): joint Newton system with one shared piece time"""
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    g = np.load(os.path.join(GOLD, "coupled.npz"))
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    wps = sc["way_points"][:2] + sc["way_points"][4:6]
    o = oa.get(kind)
    o.setup(oa.Params(4, uav_num=4, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in wps]
    tol = 1e-13 if kind == "ref" else 1e-9
    for i in range(1, 7):
        sts = o.optimization_multi(sts, coupled=True)
        for u, s in enumerate(sts):
            for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
                assert np.max(np.abs(s[k] - g["it%d_u%d_%s" % (i, u, k)])) < tol, (i, u, k)
        assert abs(sts[0]["piece_time"] - float(g["it%d_piece_time" % i])) < tol
        assert abs(sts[0]["gnorm"] - float(g["it%d_gnorm" % i])) < 1e-9 * max(1.0, float(g["it%d_gnorm" % i]))
    o.setup(oa.Params(4, ks=sc["ks"]))


def test_hostsim_tables_bit_exact(hostsim, gs):
    P = 4
    b = np.zeros((P * 8, 36)); w = np.zeros(P * 8); cv = np.zeros((P, 36)); md = np.zeros(36); kd = np.zeros(147)
    hostsim.hs_make_tables(P, 8, D(b), D(w), D(cv), D(md), D(kd))
    for k, v in dict(basis=b, weight=w, convert=cv, mdyn=md, kdop=kd).items():
        assert np.array_equal(v, gs["tab_" + k]), k


def test_hostsim_point_primitives_bit_exact(hostsim, gs):
    kd = np.ascontiguousarray(gs["tab_kdop"])
    n = len(gs["pp_P"])
    assert n > 100 and gs["pp_ok"].sum() > 20
    for i in range(n):
        Pm = np.asfortranarray(gs["pp_P"][i]); q = np.ascontiguousarray(gs["pp_q"][i])
        r = hostsim.hs_kdop_dcd(D(Pm), D(q), D(kd), C.c_double(0.2))
        assert r in (0, 3) and (r == 3) == bool(gs["pp_kdop"][i])
        v = np.zeros(3)
        hostsim.hs_gjk(D(Pm), 6, D(q), 1, D(v))
        assert np.array_equal(v, gs["pp_gjk"][i])
        c = np.zeros(3); d = C.c_double(0)
        ok = hostsim.hs_plane_point(D(Pm), D(q), C.c_double(0.2), C.c_double(0.1), D(c), C.byref(d))
        assert bool(ok) == bool(gs["pp_ok"][i])
        if ok:
            assert np.array_equal(c, gs["pp_c"][i]) and d.value == float(gs["pp_d"][i])


def test_hostsim_hull_primitives_bit_exact(hostsim, gm):
    n = len(gm["hh_P0"])
    for i in range(n):
        P0 = np.asfortranarray(gm["hh_P0"][i]); P1 = np.asfortranarray(gm["hh_P1"][i])
        c = np.zeros(3); d = C.c_double(0)
        ok = hostsim.hs_plane_hulls(D(P0), D(P1), C.c_double(0.3), D(c), C.byref(d))
        assert bool(ok) == bool(gm["hh_ok"][i])
        if ok:
            assert np.array_equal(c, gm["hh_c"][i]) and d.value == float(gm["hh_d"][i])
            d2 = C.c_double(d.value)
            hostsim.hs_refine_d(D(P0), D(P1), D(c), C.c_double(0.1), C.c_double(0.1), C.byref(d2))
            assert d2.value == float(gm["hh_dref"][i])


def test_hostsim_random_hulls_against_reference(hostsim, oracle_ref):
    """wider sweep than the fixtures: random hull pairs incl. the 12-point swept hulls of the CCD ladder"""
    o = oracle_ref
    o.setup(oa.Params(4))
    kd = np.ascontiguousarray(o.tables()["kdop"])
    rng = np.random.default_rng(3)
    for _ in range(300):
        c0 = rng.uniform(-1, 1, 3); dv = rng.normal(size=3); dv /= np.linalg.norm(dv)
        P0 = np.asfortranarray(c0 + rng.uniform(-0.3, 0.3, (6, 3)))
        P1 = np.asfortranarray(c0 + dv * rng.uniform(0.5, 1.2) + rng.uniform(-0.3, 0.3, (6, 3)))
        D0 = np.asfortranarray(rng.normal(size=(6, 3)) * 0.3); D1 = np.asfortranarray(rng.normal(size=(6, 3)) * 0.3)
        A = np.asfortranarray(np.vstack([P0, P0 + 0.7 * D0])); B = np.asfortranarray(np.vstack([P1, P1 + 0.4 * D1]))
        v = np.zeros(3)
        hostsim.hs_gjk(D(A), 12, D(B), 12, D(v))
        assert np.array_equal(v, o.gjk(A, B))
        q = np.ascontiguousarray(P1[0])
        for s in (1.0, 0.8, 0.512):
            assert bool(hostsim.hs_kdop_ccd(D(P0), D(D0), D(q), D(kd), C.c_double(0.1), C.c_double(0), C.c_double(s))) == \
                o.kdop_ccd(P0, D0, q.reshape(1, 3), 0.1, 0, s)
            assert bool(hostsim.hs_gjk_ccd(D(P0), D(D0), D(q), C.c_double(0.1), C.c_double(0), C.c_double(s))) == \
                o.gjk_ccd(P0, D0, q.reshape(1, 3), 0.1, 0, s)
            assert bool(hostsim.hs_self_gjk_ccd(D(P0), D(D0), D(P1), D(D1), C.c_double(0.1), C.c_double(s), C.c_double(0.8 * s))) == \
                o.self_gjk_ccd(P0, D0, P1, D1, 0.1, 0, s, 0, 0.8 * s)


# ---------------------------------------------------------------------------------------------------------------
def test_scene_init_matches_reference_layout(gs):
    sp = scenes.init_spline_single(gs["way_points"])
    assert np.array_equal(sp, gs["it0_spline"])
    st = scenes.init_state(sp)
    assert np.array_equal(st["p_slack"], gs["it0_p_slack"])


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "trajopt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(tob_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 30
    so = os.path.join(ROOT, "traj-opt-admm_b200", "libtrajopt_b200.so")
    assert os.path.exists(so), "library not built: run python -c 'import __graft_entry__ as g; g.build()'"
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = set(re.findall(r" T (tob_[a-z0-9_]+)", syms))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    lib = C.CDLL(so)   # loads without a GPU
    for n in names:
        getattr(lib, n)


def test_library_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from trajopt import api
    with pytest.raises(RuntimeError) as ei:
        api.Solver(4)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_hostsim_kdop_filter_equals_fp64_gate(hostsim, gs):
    """The single-precision filter of k_narrow's 49-DOP gate (csrc/gjk.cuh: kdop_point_gate, thresholds as made by
    segments.cu) must take the decision of the FP64 gate (CCD::KDOPDCD, CCD.h:354-413) for every point: random points
    around the hull, points placed ON each threshold (lv = lo - d, lv = hi + d) and one / a few ulps to either side,
    large coordinate offsets, tiny hulls."""
    kd = np.ascontiguousarray(gs["tab_kdop"])
    rng = np.random.Generator(np.random.PCG64(77))
    hostsim.hs_kdop_gate_batch.restype = C.c_int
    total_exact = 0
    total = 0
    axes = kd.reshape(49, 3)
    for trial in range(60):
        scale = [1.0, 1.0, 0.05, 30.0][trial % 4]
        shift = [0.0, 7.0, 250.0, -3.0e4, 1.0e7][trial % 5]
        P = rng.normal(size=(6, 3)) * 0.4 * scale + rng.normal(size=3) * 3 + shift
        d = [0.2, 0.05, 1e-3, 0.0][trial % 4] * scale
        lo = (P @ axes.T).min(axis=0); hi = (P @ axes.T).max(axis=0)
        pts = [P.mean(axis=0) + rng.normal(size=(3000, 3)) * (0.5 * scale + d)]
        # points on the thresholds: start from a point near the hull, move it along axis k until its level hits the threshold
        base = P.mean(axis=0) + rng.normal(size=(49, 3)) * 0.1 * scale
        for sgn, thr in ((-1.0, lo - d), (1.0, hi + d)):
            lv = np.einsum("ij,ij->i", base, axes)
            on = base + ((thr - lv)[:, None]) * axes
            for ulps in (0, 1, -1, 3, -3, 40, -40, 1000, -1000):
                p = on.copy()
                j = np.argmax(np.abs(axes), axis=1)
                x = p[np.arange(49), j]
                p[np.arange(49), j] = x + ulps * np.spacing(np.abs(x))
                pts.append(p)
        pts = np.ascontiguousarray(np.concatenate(pts))
        n = len(pts)
        of = np.zeros(n, dtype=np.uint8); ox = np.zeros(n, dtype=np.uint8); ne = C.c_ulonglong(0)
        hostsim.hs_kdop_gate_batch(D(np.asfortranarray(P)), D(pts), n, D(kd), C.c_double(d),
                                   of.ctypes.data_as(C.POINTER(C.c_ubyte)), ox.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(ne))
        assert np.array_equal(of, ox), (trial, int((of != ox).sum()))
        assert 0 < ox[:3000].sum() < 3000          # the random part exercises both outcomes
        if shift == 0.0 or abs(shift) < 10:
            total_exact += ne.value; total += 3000
    assert total_exact > 0                         # the threshold points reach the FP64 fallback


def test_fast_log(hostsim):
    """csrc/fastlog.cuh (the logarithm of the barrier kernels) against the host's long-double logarithm: at most 2 ulp over the
    band 0 < x <= 1 the barrier uses (random, log-uniform down to 1e-12, next to 1, every bin edge), a few ulp above 1, the
    library's answers for everything that is not a positive normal number; the bin that holds 1.0 is the exact one."""
    import struct
    rng = np.random.default_rng(0)

    def ulps(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = len(x)
        o, hi, lo = np.empty(n), np.empty(n), np.empty(n)
        hostsim.hs_fast_log(D(x), n, D(o))
        hostsim.hs_ref_logl(D(x), n, D(hi), D(lo))
        with np.errstate(invalid="ignore"):
            return np.abs((o - hi) - lo) / np.maximum(np.spacing(np.abs(hi)), 5e-324), o

    edges = []
    for e in (-40, -3, -1, 0):
        for i in range(129):
            h = 0x3FE6AAAB + (i << 13)
            for d in (-1, 0, 1):
                v = struct.unpack("<d", struct.pack("<Q", ((h + d) << 32) | (0xFFFFFFFF if d < 0 else 0)))[0]
                edges.append(v * 2.0 ** e)
    edges = np.array(edges)
    for x in (rng.uniform(0, 1, 400_000), 10 ** rng.uniform(-12, 0, 400_000), 1 - 10 ** rng.uniform(-16, -2, 200_000),
              edges[edges <= 1.0]):
        u, _ = ulps(x)
        assert u.max() <= 2.0, u.max()
    for x in (10 ** rng.uniform(-300, 300, 200_000), 1 + 10 ** rng.uniform(-16, -1, 200_000), edges):
        u, _ = ulps(x)
        assert u.max() <= 6.0, u.max()
    _, o = ulps(np.array([1.0, 0.0, -1.0, np.inf, np.nan, 5e-324, 2.0 ** -1022]))
    assert o[0] == 0.0 and o[1] == -np.inf and np.isnan(o[2]) and o[3] == np.inf and np.isnan(o[4])
    assert o[5] == np.log(5e-324) and abs(o[6] - np.log(2.0 ** -1022)) < 1e-12


def test_hostsim_narrow_rule(hostsim, gs):
    """k_narrow's decision in the product's own arithmetic (host build of gjk.cuh): the first 14 axes of the gate, GJK, the
    other 35 axes only for a witness within 1e-6 of the gap -- against the reference's order (all 49 axes, then GJK): the same
    pairs get a plane, with the same bits, and less than half of the axis groups is evaluated.  Rows: the hulls of the golden
    point-primitive vectors; points: a cloud around each hull, plus points placed at the gap itself (inside the band, where the
    rest of the gate does run)."""
    kd = np.ascontiguousarray(gs["tab_kdop"], dtype=np.float64)
    rng = np.random.default_rng(11)
    dist, offset = 0.2, 0.1
    tot_planes = tot_band = g_ref = g_cut = 0
    for r in range(0, len(gs["pp_P"]), max(1, len(gs["pp_P"]) // 12)):
        Pm = np.asfortranarray(gs["pp_P"][r])
        centre = Pm.mean(axis=0)
        pts = centre + rng.normal(0, 0.22, size=(4000, 3))
        # points at the gap: witness length = dist up to rounding, some inside the band
        dirs = rng.normal(size=(200, 3)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        far = Pm[np.argmax(np.linalg.norm(Pm - centre, axis=1))]          # a vertex of the hull, pushed outwards
        out = (far - centre) / np.linalg.norm(far - centre)
        dirs = dirs + 2.0 * out; dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        edge = far + dirs * (dist * (1 - rng.uniform(0, 2e-6, size=(200, 1))))
        pts = np.ascontiguousarray(np.vstack([pts, edge]))
        n = len(pts)
        ok_r = np.zeros(n, np.uint8); ok_c = np.zeros(n, np.uint8); pl_r = np.zeros((n, 4)); pl_c = np.zeros((n, 4))
        nb, gr, gc = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
        hostsim.hs_narrow_rule(D(Pm), D(pts), n, D(kd), C.c_double(dist), C.c_double(offset), 14,
                               ok_r.ctypes.data_as(C.POINTER(C.c_ubyte)), D(pl_r), ok_c.ctypes.data_as(C.POINTER(C.c_ubyte)), D(pl_c),
                               C.byref(nb), C.byref(gr), C.byref(gc))
        assert np.array_equal(ok_r, ok_c)
        assert np.array_equal(pl_r.view(np.uint64), pl_c.view(np.uint64))
        tot_planes += int(ok_r.sum()); tot_band += nb.value; g_ref += gr.value; g_cut += gc.value
    assert tot_planes > 5000 and tot_band > 20, (tot_planes, tot_band)
    assert g_cut < 0.5 * g_ref
