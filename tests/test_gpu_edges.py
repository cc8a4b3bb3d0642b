"""Edge cases of the hot path through the C ABI: nothing near the trajectory, a cloud smaller than one LBVH leaf, candidate
buffers that overflow in the middle of a device-resident iteration (grow + repeat must not change the result)."""
import os

import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu
KEYS = ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda")


def test_no_candidates_at_all(oracle_any):
    """every obstacle is far away: empty candidate / plane lists, the iteration is pure smoothing and matches the oracle"""
    sc = scenes.bridge(n_pts=4000, seed=4)
    V = sc["V"] + np.array([0.0, 50.0, 0.0])
    P = len(sc["way_points"][0]) - 1
    o = oracle_any
    o.setup(oa.Params(P, ks=sc["ks"])); o.init_pointcloud(V)
    s = api.Solver(P, ks=sc["ks"]); s.init_pointcloud(V)
    st = scenes.initial_states(sc)[0]
    off, ids = s.dcd_collision(st["spline"], 0.2)
    assert len(ids) == 0 and not off.any()
    po, pc, pd = s.separate_planes([st["spline"]])
    assert len(pd) == 0
    a = b = st
    for _ in range(3):
        a, b = o.optimization(a), s.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6
    assert s.counters()["planes"] == 0


@pytest.mark.parametrize("n", [1, 7, 33])
def test_cloud_smaller_than_a_leaf(oracle_any, n):
    """1, 7 and 33 points (a leaf holds 32): padded leaves never produce candidates, real points do"""
    sc = scenes.bridge(n_pts=4000, seed=4)
    P = len(sc["way_points"][0]) - 1
    st = scenes.initial_states(sc)[0]
    rng = np.random.default_rng(n)
    t = rng.uniform(0.1, 0.9, n)
    wp = np.asarray(sc["way_points"][0])
    V = wp[0] + np.outer(t, wp[-1] - wp[0]) + np.array([0.0, 0.0, 0.16]) + rng.normal(scale=0.01, size=(n, 3))
    o = oracle_any
    o.setup(oa.Params(P, ks=sc["ks"])); o.init_pointcloud(V)
    s = api.Solver(P, ks=sc["ks"]); s.init_pointcloud(V)
    ro, ri = o.dcd_collision(st["spline"], 0.2)
    go, gi = s.dcd_collision(st["spline"], 0.2)
    assert np.array_equal(ro, go) and len(gi) >= n
    for r in range(len(ro) - 1):
        assert np.array_equal(np.sort(ri[ro[r]:ro[r + 1]]), gi[go[r]:go[r + 1]])
    a = b = st
    for _ in range(2):
        a, b = o.optimization(a), s.optimization(b)
    assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6


def test_candidate_overflow_in_a_resident_iteration():
    """the candidate buffers start far too small: the device-side guard leaves the state untouched, the host grows the buffers
    and repeats the iteration; bitwise the same as with buffers that were large enough from the start"""
    sc = scenes.bridge(n_pts=20000, seed=21)
    P = len(sc["way_points"][0]) - 1
    st = scenes.initial_states(sc)
    res = []
    for cap in (None, 256):
        if cap:
            os.environ["TRAJOPT_B200_CAND_CAP"] = str(cap)
        try:
            s = api.Solver(P, ks=sc["ks"])
            s.init_pointcloud(sc["V"])
            s.states_upload(st)
            for _ in range(3):
                s.iterate(1)
            res.append((s.states_download(st)[0], s.counters()["dcd_candidates"]))
        finally:
            os.environ.pop("TRAJOPT_B200_CAND_CAP", None)
    (a, ca), (b, cb) = res
    assert ca == cb and ca > 3 * 256
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["piece_time"] == b["piece_time"]
