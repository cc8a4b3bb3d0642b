"""The result of an iteration must not depend on HOW the line search is scheduled: trial points per launch, rounds launched
ahead of the host, one call of several iterations or several calls of one.  (The multi-GPU sharded run picks another
schedule than the single context because it has fewer rows per context; tests/run_sharded_check.py compares the two
bitwise on 2 GPUs.  A rung laid out by k_robot_ls once got a trial time one ulp away from the same rung laid out by
k_ls_init: one file is built with FMA contraction, the other without.)"""
import os

import numpy as np
import pytest

from trajopt import api, scenes

pytestmark = pytest.mark.gpu
KEYS = ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda")


def run(sc, sts, policy, iters, uav_num=1, one_call=False):
    if policy:
        os.environ["TRAJOPT_B200_LS"] = policy
    else:
        os.environ.pop("TRAJOPT_B200_LS", None)
    try:
        P = len(sc["way_points"][0]) - 1
        s = api.Solver(P, uav_num=uav_num, ks=sc["ks"])
        s.init_pointcloud(sc["V"])
        s.states_upload(sts)
        if one_call:
            s.iterate(iters)
        else:
            for _ in range(iters):
                s.iterate(1)
        return s.states_download(sts)
    finally:
        os.environ.pop("TRAJOPT_B200_LS", None)


def same(a, b):
    for x, y in zip(a, b):
        for k in KEYS:
            if not np.array_equal(x[k], y[k]):
                return False
        if x["piece_time"] != y["piece_time"]:
            return False
    return True


def test_single_uav_result_independent_of_line_search_schedule():
    sc = scenes.bridge(n_pts=8000, seed=3)
    sts = scenes.initial_states(sc)
    ref = run(sc, sts, None, 10)
    for policy in ("3,9,3", "9,9,2", "2,3,8", "5,5,4"):
        assert same(ref, run(sc, sts, policy, 10)), policy
    assert same(ref, run(sc, sts, None, 10, one_call=True))
    # the ladder with every rung evaluated (default: the rungs that violate a bound for certain are skipped, k_ls_bound_mask)
    for skip in ("0",):
        os.environ["TRAJOPT_B200_LS_SKIP"] = skip
        try:
            assert same(ref, run(sc, sts, None, 10)), skip
            assert same(ref, run(sc, sts, "2,3,8", 10)), skip
        finally:
            os.environ.pop("TRAJOPT_B200_LS_SKIP", None)


def test_multi_uav_result_independent_of_line_search_schedule():
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in sc["way_points"]]
    ref = run(sc, sts, None, 6, uav_num=len(sts))
    for policy in ("3,9,3", "9,9,2", "2,3,8"):
        assert same(ref, run(sc, sts, policy, 6, uav_num=len(sts))), policy
    for skip in ("0",):
        os.environ["TRAJOPT_B200_LS_SKIP"] = skip
        try:
            assert same(ref, run(sc, sts, None, 6, uav_num=len(sts))), skip
        finally:
            os.environ.pop("TRAJOPT_B200_LS_SKIP", None)


# ---- many rows (>= 8192: the throughput regime picks other kernel variants and another line-search policy) ------------------
def _batch(n=130):
    rng = np.random.Generator(np.random.PCG64(5))
    return [scenes.tube(int(rng.integers(1200, 4000)), 100 + i, float(rng.uniform(0.14, 0.5))) for i in range(n)]


def _run_batch(scs, sts, iters, env):
    keys = ("TRAJOPT_B200_LS", "TRAJOPT_B200_EN_OCC", "TRAJOPT_B200_NP_FILTER", "TRAJOPT_B200_CCD_OCC", "TRAJOPT_B200_NP_OCC",
            "TRAJOPT_B200_PACK_GRID", "TRAJOPT_B200_NP_BAND", "TRAJOPT_B200_NP_GATE1", "TRAJOPT_B200_NP_PMEM", "TRAJOPT_B200_LS_SKIP", "TRAJOPT_B200_PIECE_CTA",
            "TRAJOPT_B200_NP_CHUNK")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        s = api.Solver(8, uav_num=len(scs), ks=1e-8)
        s.init_pointclouds([sc["V"] for sc in scs])
        s.states_upload(sts)
        for _ in range(iters):
            s.iterate(1, mode=2)
        return s.states_download(sts), s.counters()
    finally:
        for k in keys:
            os.environ.pop(k, None)


def test_many_rows_result_independent_of_kernel_variants_and_schedule():
    """130 independent problems = 8320 rows.  The default run (occupancy variants of k_row_energy / k_bp_ccd, the
    single-precision filter of the 49-DOP gate, the many-row line-search policy) must be bitwise equal to the run with every
    variant switched off, to other line-search schedules, and -- problem by problem -- to a single-problem context."""
    scs = _batch()
    sts = [scenes.initial_states(sc)[0] for sc in scs]
    ref, cref = _run_batch(scs, sts, 5, {})
    assert cref["np_kdop_groups"] > 0
    for env in ({"TRAJOPT_B200_EN_OCC": "0", "TRAJOPT_B200_NP_FILTER": "0", "TRAJOPT_B200_CCD_OCC": "4"},
                {"TRAJOPT_B200_EN_OCC": "2", "TRAJOPT_B200_LS": "2,5,5", "TRAJOPT_B200_CCD_OCC": "12"},
                {"TRAJOPT_B200_LS": "2,2,16", "TRAJOPT_B200_PACK_GRID": "4"}, {"TRAJOPT_B200_LS": "2,3,9"},
                {"TRAJOPT_B200_LS": "9,9,2", "TRAJOPT_B200_NP_OCC": "6"},
                {"TRAJOPT_B200_NP_BAND": "0"}, {"TRAJOPT_B200_NP_BAND": "0", "TRAJOPT_B200_NP_FILTER": "0", "TRAJOPT_B200_NP_GATE1": "49"},
                {"TRAJOPT_B200_NP_GATE1": "7"}, {"TRAJOPT_B200_NP_PMEM": "5"}, {"TRAJOPT_B200_NP_PMEM": "4"},
                {"TRAJOPT_B200_LS_SKIP": "0"}, {"TRAJOPT_B200_LS_SKIP": "0", "TRAJOPT_B200_LS": "2,5,5,2"}, {"TRAJOPT_B200_LS": "2,3,6,2"},
                {"TRAJOPT_B200_LS": "2,5,3,3"}, {"TRAJOPT_B200_PIECE_CTA": "128", "TRAJOPT_B200_NP_CHUNK": "256"},
                {"TRAJOPT_B200_PIECE_CTA": "384", "TRAJOPT_B200_NP_CHUNK": "128"}):
        got, cgot = _run_batch(scs, sts, 5, env)
        assert same(ref, got), env
        assert cgot["planes"] == cref["planes"] and cgot["dcd_candidates"] == cref["dcd_candidates"], env
        if env.get("TRAJOPT_B200_NP_FILTER") == "0":
            assert cgot["np_kdop_exact"] == 0
            if "TRAJOPT_B200_NP_GATE1" not in env:
                assert cgot["np_kdop_groups"] == cref["np_kdop_groups"]
        if env.get("TRAJOPT_B200_NP_BAND") == "0":      # every accepted pair went through the rest of the gate
            assert cgot["np_band"] >= cgot["planes"] > 0
    for u in (0, 57, 129):
        s1 = api.Solver(8, uav_num=1, ks=1e-8)
        s1.init_pointcloud(scs[u]["V"])
        s1.states_upload([sts[u]])
        s1.iterate(5)
        assert same([ref[u]], s1.states_download([sts[u]])), u


def test_kdop_filter_same_planes_as_fp64_gate():
    """plane sets (offsets, c, d) with the single-precision gate filter on and off: bit-identical (forest-like scene, 64 pieces)"""
    sc = scenes.forest(n_pts=200_000, seed=2)
    st = scenes.initial_states(sc)[0]
    P = len(sc["way_points"][0]) - 1
    out = []
    for filt in ("1", "0"):
        os.environ["TRAJOPT_B200_NP_FILTER"] = filt
        try:
            s = api.Solver(P, ks=sc["ks"])
            s.init_pointcloud(sc["V"])
            out.append(s.separate_plane(st["spline"]) + (s.counters(),))
        finally:
            os.environ.pop("TRAJOPT_B200_NP_FILTER", None)
    (o1, c1, d1, k1), (o0, c0, d0, k0) = out
    assert len(d1) > 1000
    assert np.array_equal(o1, o0) and np.array_equal(c1, c0) and np.array_equal(d1, d0)
    assert k0["np_kdop_exact"] == 0 and k1["np_kdop_groups"] == k0["np_kdop_groups"]
    assert k1["np_kdop_exact"] < 1e-3 * 7 * k1["np_kdop_groups"]        # the fallback is rare


def test_gate_implied_by_gjk_distance_same_planes_as_full_gate():
    """Plane sets with the 49-DOP gate cut short by the GJK distance (default: axes 15..49 only for pairs within 1e-6 of the
    gap) against the reference order (all 49 axes first, then GJK: TRAJOPT_B200_NP_GATE1=49) and against the forced band
    (rest of the gate for every accepted pair): bit-identical offsets and planes; the band itself is nearly empty."""
    sc = scenes.forest(n_pts=200_000, seed=2)
    st = scenes.initial_states(sc)[0]
    P = len(sc["way_points"][0]) - 1
    out = []
    keys = ("TRAJOPT_B200_NP_BAND", "TRAJOPT_B200_NP_GATE1", "TRAJOPT_B200_NP_FILTER", "TRAJOPT_B200_NP_PMEM")
    for env in ({}, {"TRAJOPT_B200_NP_GATE1": "49", "TRAJOPT_B200_NP_FILTER": "0"}, {"TRAJOPT_B200_NP_BAND": "0"},
                {"TRAJOPT_B200_NP_GATE1": "7"}, {"TRAJOPT_B200_NP_GATE1": "28"}, {"TRAJOPT_B200_NP_PMEM": "6"}):
        os.environ.update(env)
        try:
            s = api.Solver(P, ks=sc["ks"])
            s.init_pointcloud(sc["V"])
            out.append(s.separate_plane(st["spline"]) + (s.counters(),))
        finally:
            for k in keys:
                os.environ.pop(k, None)
    o0, c0, d0, k0 = out[0]
    assert len(d0) > 1000
    for o, c, d, k in out[1:]:
        assert np.array_equal(o, o0) and np.array_equal(c, c0) and np.array_equal(d, d0)
        assert k["planes"] == k0["planes"]
    assert k0["np_band"] <= 1e-4 * k0["planes"] + 2
    assert out[1][3]["np_band"] <= 1e-4 * k0["planes"] + 2 and out[2][3]["np_band"] >= k0["planes"]
    assert k0["np_kdop_groups"] < out[1][3]["np_kdop_groups"]       # the point of it: fewer axes evaluated


def test_broadphase_item_records_same_candidates():
    """The fill pass scatters from the item records of the count pass (default), repeats the tests (TRAJOPT_B200_BP_REC=0), or
    does either per CTA when the records run out of room (a capacity of 300 records): identical candidate lists and planes."""
    sc = scenes.forest(n_pts=200_000, seed=2)
    st = scenes.initial_states(sc)[0]
    P = len(sc["way_points"][0]) - 1
    out = []
    for rec in (None, "0", "300"):
        if rec is not None:
            os.environ["TRAJOPT_B200_BP_REC"] = rec
        try:
            s = api.Solver(P, ks=sc["ks"])
            s.init_pointcloud(sc["V"])
            out.append(s.dcd_collision(st["spline"], 0.2) + s.ccd_collision(st["spline"], 0.05 * np.ones_like(st["spline"]), 0.1)
                       + s.separate_plane(st["spline"]))
        finally:
            os.environ.pop("TRAJOPT_B200_BP_REC", None)
    assert len(out[0][1]) > 10_000
    for o in out[1:]:
        assert len(o) == len(out[0])
        for x, y in zip(out[0], o):
            assert np.array_equal(x, y)
