"""GPU parity of the batched independent-problem mode (BASELINE config 5): every robot slot of one context holds its own
single-UAV problem with its own cloud; one launch sequence iterates all of them.  Each problem must match the oracle run
on that problem alone (Optimization3D_admm::optimization, Optimization3D_admm.h:29-67)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def members():
    from trajopt import scenes
    return [scenes.tube(9000, 11, 0.16), scenes.tube(30011, 12, 0.45), scenes.tube(1500, 13, 0.19), scenes.bridge(20000, seed=5)]


def test_batch_broadphase_and_planes_bit_exact(oracle_any):
    from trajopt import api, scenes
    from oracle import oracle_api as oa
    scs = members()
    P = 8
    s = api.Solver(P, uav_num=len(scs), ks=1e-8)
    s.init_pointclouds([sc["V"] for sc in scs])
    sts = [scenes.initial_states(sc)[0] for sc in scs]
    off, ids = s.dcd_collision([st["spline"] for st in sts], 0.2)
    poff, pc, pd = s.separate_planes([st["spline"] for st in sts], False)
    n_tr = P * 8
    o = oracle_any
    for u, sc in enumerate(scs):
        o.setup(oa.Params(P, ks=1e-8)); o.init_pointcloud(sc["V"])
        ro, ri = o.dcd_collision(sts[u]["spline"], 0.2)
        go = off[u * n_tr:(u + 1) * n_tr + 1]
        assert np.array_equal(go - go[0], ro), u
        for r in range(n_tr):
            assert np.array_equal(np.sort(ri[ro[r]:ro[r + 1]]), ids[go[r]:go[r + 1]]), (u, r)
        qo, qc, qd = o.separate_plane(sts[u]["spline"])
        mo = poff[u * n_tr:(u + 1) * n_tr + 1]
        assert np.array_equal(mo - mo[0], qo), u
        for r in range(n_tr):
            a = np.column_stack([qc[qo[r]:qo[r + 1]], qd[qo[r]:qo[r + 1]]])
            b = np.column_stack([pc[mo[r]:mo[r + 1]], pd[mo[r]:mo[r + 1]]])
            a = a[np.lexsort(a.T[::-1])]; b = b[np.lexsort(b.T[::-1])]
            assert np.array_equal(a, b), (u, r)


def test_batch_iterations_track_the_oracle_per_problem(oracle_any):
    from trajopt import api, scenes
    from oracle import oracle_api as oa
    scs = members()
    P = 8
    s = api.Solver(P, uav_num=len(scs), ks=1e-8)
    s.init_pointclouds([sc["V"] for sc in scs])
    sts = [scenes.initial_states(sc)[0] for sc in scs]
    s.states_upload(sts)
    for _ in range(3):
        s.iterate(1, mode=2)
    got = s.states_download(sts)
    o = oracle_any
    for u, sc in enumerate(scs):
        o.setup(oa.Params(P, ks=1e-8)); o.init_pointcloud(sc["V"])
        ref = sts[u]
        for _ in range(3):
            ref = o.optimization(ref)
        assert np.max(np.abs(ref["spline"] - got[u]["spline"])) < 1e-6, u
        assert abs(ref["piece_time"] - got[u]["piece_time"]) < 1e-6, u
        assert np.max(np.abs(ref["p_slack"] - got[u]["p_slack"])) < 1e-6, u
    # the same problems one by one through single-robot contexts: bitwise identical to the batched run
    for u, sc in enumerate(scs):
        s1 = api.Solver(P, uav_num=1, ks=1e-8)
        s1.init_pointcloud(sc["V"])
        s1.states_upload([sts[u]])
        s1.iterate(3)
        one = s1.states_download([sts[u]])[0]
        assert np.array_equal(one["spline"], got[u]["spline"]), u
        assert one["piece_time"] == got[u]["piece_time"], u
