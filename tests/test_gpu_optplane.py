"""GPU parity of the persistent-plane mode ("optimal_plane": 1; SURVEY.md section 8a rows P4 / P5 and the is_optimal_plane
branches of P6): the CUDA path through the C ABI against the oracle and against tests/golden/optplane.npz (produced by the
unmodified reference).

Tolerances: the refinement uses sin / cos / log, so (c, d) are tolerance-matched: 1e-9 for optimal_cd (well conditioned);
self_optimal_cd as in tests/test_cpu_optplane.py (1e-6 on the pairs the golden file marks stable, algorithmic properties on
all).  Live-plane SETS (which (sub-segment, point) pairs own a plane) are exact; trajectories 1e-6."""
import os

import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa
from test_cpu_optplane import check_self_pairs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def go():
    return np.load(os.path.join(ROOT, "tests", "golden", "optplane.npz"))


def test_optimal_cd_batch_against_golden(go):
    s = api.Solver(4, ks=1e-8, optimal_plane=1)
    c, d, capped = s.optimal_cd_batch(go["cd_P"], go["cd_q"], go["cd_c0"], go["cd_d0"])
    assert not capped.any()
    assert np.abs(c - go["cd_c1"]).max() < 1e-9 and np.abs(d - go["cd_d1"]).max() < 1e-9
    assert np.abs(go["cd_c1"] - go["cd_c0"]).max() > 1e-3      # the refinement really moved the planes


def test_self_optimal_cd_batch_against_golden(go):
    s = api.Solver(4, uav_num=4, ks=1e-3, optimal_plane=1)
    c, d, capped = s.self_optimal_cd_batch(go["scd_P0"], go["scd_P1"], go["scd_c0"], go["scd_d0"])
    assert not capped.any()
    it = iter(range(len(d)))
    check_self_pairs(go, lambda P0, P1, c0, d0: (lambda i: (c[i], float(d[i])))(next(it)))


def live_sorted(rows, ids, c, d):
    k = np.lexsort((ids, rows))
    return rows[k], ids[k], c[k], d[k]


def test_persistent_iterations_against_golden(go):
    """6 ADMM iterations of the golden single-UAV scene: same live (sub-segment, point) pairs, planes and trajectory"""
    sc = scenes.bridge(n_pts=4000, seed=21, n_pieces=4)
    s = api.Solver(4, ks=sc["ks"], optimal_plane=1)
    s.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    for i in range(1, 7):
        st = s.optimization(st)
        assert np.abs(st["spline"] - go["s_it%d_spline" % i]).max() < 1e-6, i
        assert abs(st["piece_time"] - float(go["s_it%d_piece_time" % i])) < 1e-6
        if i in (1, 3, 6):
            rows, ids, c, d = live_sorted(*s.live_planes())
            assert np.array_equal(rows, go["s_it%d_live_tr" % i]) and np.array_equal(ids, go["s_it%d_live_id" % i]), i
            assert np.abs(c - go["s_it%d_live_c" % i]).max() < 1e-8 and np.abs(d - go["s_it%d_live_d" % i]).max() < 1e-8, i
    ctr = s.counters()
    assert ctr["live_planes"] == len(go["s_it6_live_d"]) and ctr["refine_capped"] == 0
    # planes_reset empties the set: the next plane pass starts from the GJK planes again
    s.planes_reset()
    assert len(s.live_planes()[0]) == 0


def test_persistent_iterations_against_oracle(oracle_any):
    """a larger cloud, resident iterations (CUDA graph path) vs the oracle run with is_optimal_plane"""
    sc = scenes.bridge(n_pts=8000, seed=3)
    P = len(sc["way_points"][0]) - 1
    o = oracle_any
    o.setup(oa.Params(P, ks=sc["ks"], optimal_plane=1)); o.init_pointcloud(sc["V"]); o.reset_persistent_planes()
    s = api.Solver(P, ks=sc["ks"], optimal_plane=1)
    s.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    s.states_upload([st])
    ref = st
    for i in range(6):
        ref = o.optimization(ref)
        s.iterate(1)
        got = s.states_download([st])[0]
        assert np.abs(ref["spline"] - got["spline"]).max() < 1e-6, i
        tr, ids, c, d = o.live_planes()
        rows, gids, gc, gd = live_sorted(*s.live_planes())
        assert np.array_equal(tr, rows) and np.array_equal(ids, gids), i
        assert np.abs(c - gc).max() < 1e-8 and np.abs(d - gd).max() < 1e-8, i
    assert len(tr) > 1500
    o.setup(oa.Params(P, ks=sc["ks"]))


def test_persistent_live_set_growth_preserves_planes():
    """the live set outgrows its buffer: the overflow path re-allocates, keeps the planes and repeats the iteration"""
    sc = scenes.bridge(n_pts=8000, seed=3)
    P = len(sc["way_points"][0]) - 1
    st = scenes.initial_states(sc)[0]
    res = []
    for cap in (None, 64):
        s = api.Solver(P, ks=sc["ks"], optimal_plane=1)
        s.init_pointcloud(sc["V"])
        if cap:
            os.environ["TRAJOPT_B200_LIVE_CAP"] = str(cap)
        try:
            s.states_upload([st])
            s.iterate(3)
        finally:
            os.environ.pop("TRAJOPT_B200_LIVE_CAP", None)
        res.append((s.states_download([st])[0]["spline"], s.live_planes()))
    assert np.array_equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)


def test_persistent_batch_equals_single_contexts():
    """mode 2 (independent problems) with persistent planes: every problem bitwise equal to its own single-UAV context"""
    scs = [scenes.tube(9000, 11, 0.16), scenes.tube(1500, 13, 0.19), scenes.bridge(8000, seed=5)]
    P = 8
    s = api.Solver(P, uav_num=len(scs), ks=1e-8, optimal_plane=1)
    s.init_pointclouds([sc["V"] for sc in scs])
    sts = [scenes.initial_states(sc)[0] for sc in scs]
    s.states_upload(sts)
    for _ in range(3):
        s.iterate(1, mode=2)
    got = s.states_download(sts)
    rows, ids, c, d = s.live_planes()
    for u, sc in enumerate(scs):
        s1 = api.Solver(P, uav_num=1, ks=1e-8, optimal_plane=1)
        s1.init_pointcloud(sc["V"])
        s1.states_upload([sts[u]])
        s1.iterate(3)
        one = s1.states_download([sts[u]])[0]
        assert np.array_equal(one["spline"], got[u]["spline"]), u
        r1, i1, c1, d1 = s1.live_planes()
        m = (rows >= u * P * 8) & (rows < (u + 1) * P * 8)
        assert np.array_equal(r1 + u * P * 8, rows[m]) and np.array_equal(c1, c[m]) and np.array_equal(d1, d[m]), u


def test_persistent_multi_against_golden(go):
    """decoupled 4-UAV iterations with persistent inter-robot planes (self_optimal_cd every iteration)"""
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    wps = sc["way_points"][:2] + sc["way_points"][4:6]
    s = api.Solver(4, uav_num=4, ks=sc["ks"], optimal_plane=1)
    s.init_pointcloud(sc["V"])
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in wps]
    for i in range(1, 5):
        sts = s.optimization(sts)
        for u, st in enumerate(sts):
            assert np.abs(st["spline"] - go["m_it%d_u%d_spline" % (i, u)]).max() < 1e-6, (i, u)
    assert s.counters()["refine_capped"] == 0
