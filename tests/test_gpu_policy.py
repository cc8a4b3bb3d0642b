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


def test_multi_uav_result_independent_of_line_search_schedule():
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in sc["way_points"]]
    ref = run(sc, sts, None, 6, uav_num=len(sts))
    for policy in ("3,9,3", "9,9,2", "2,3,8"):
        assert same(ref, run(sc, sts, policy, 6, uav_num=len(sts))), policy
