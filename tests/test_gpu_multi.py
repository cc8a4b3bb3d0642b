"""GPU parity tests of the multi-robot path (Optimization3D_multi::optimization_decouple) through the C ABI."""
import os

import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 8


def rows_sorted(off, c, d):
    out = []
    for r in range(len(off) - 1):
        blk = np.column_stack([c[off[r]:off[r + 1]], d[off[r]:off[r + 1]]])
        if len(blk):
            blk = blk[np.lexsort(blk.T[::-1])]
        out.append(blk)
    return out


@pytest.fixture(scope="module")
def world(oracle_any):
    sc = scenes.cross(n_pts=20000, seed=3)
    U = sc["uav_num"]
    o = oracle_any
    o.setup(oa.Params(P, uav_num=U, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, uav_num=U, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    st0 = scenes.initial_states(sc)
    st = st0
    for _ in range(3):
        st = o.optimization_multi(st, coupled=False)
    return dict(sc=sc, o=o, s=s, st0=st0, st=st, U=U)


def test_hull_primitives_against_golden(world):
    g = np.load(os.path.join(ROOT, "tests", "golden", "multi.npz"))
    s = world["s"]
    ok, c, d = s.plane_hulls_batch(g["hh_P0"], g["hh_P1"], 0.3, refine=False)
    assert np.array_equal(ok, g["hh_ok"])
    m = g["hh_ok"]
    assert np.array_equal(c[m], g["hh_c"][m]) and np.array_equal(d[m], g["hh_d"][m])
    ok2, c2, d2 = s.plane_hulls_batch(g["hh_P0"], g["hh_P1"], 0.3, refine=True)
    assert np.max(np.abs(d2[m] - g["hh_dref"][m])) <= 1e-12


@pytest.mark.parametrize("which", ["st0", "st"])
def test_planes_with_inter_robot_terms(world, which):
    o, s, U = world["o"], world["s"], world["U"]
    splines = [x["spline"] for x in world[which]]
    go, gc, gd = s.separate_planes(splines, with_self=True)
    so, sc_, sd = o.separate_self(splines)
    n_tr = P * 8
    n_self = 0
    got = rows_sorted(go, gc, gd)
    for u in range(U):
        ro, rc, rd = o.separate_plane(splines[u])
        for tr in range(n_tr):
            r = u * n_tr + tr
            ob = np.column_stack([rc[ro[tr]:ro[tr + 1]], rd[ro[tr]:ro[tr + 1]]])
            se = np.column_stack([sc_[so[r]:so[r + 1]], sd[so[r]:so[r + 1]]])
            n_self += len(se)
            ref = np.vstack([ob, se])
            assert len(ref) == len(got[r]), (u, tr)
            if len(ref):
                ref = ref[np.lexsort(ref.T[::-1])]
                # obstacle planes bit-exact; inter-robot d comes out of a Newton loop using log(): 1e-12
                assert np.max(np.abs(ref - got[r])) <= 1e-12
    assert n_self > 0


def test_self_step(world):
    o, s, U = world["o"], world["s"], world["U"]
    splines = [x["spline"] for x in world["st"]]
    rng = np.random.default_rng(11)
    seen = set()
    for trial in range(4):
        dirs = []
        for sp in splines:
            dd = np.zeros_like(sp); dd[2:-2] = rng.normal(size=(sp.shape[0] - 4, 3)) * (0.3 + 0.3 * trial)
            dirs.append(np.asfortranarray(dd))
        ref = o.self_step(splines, dirs)
        got = s.self_step(splines, dirs)
        assert np.all(got <= ref)
        assert np.array_equal(got, ref)
        seen.update(ref.tolist())
        assert s.self_step(splines, dirs, coupled=True) == o.couple_self_step(splines, dirs)
    assert len(seen) > 1


def test_decoupled_iterations_track_reference(world):
    o, s, U = world["o"], world["s"], world["U"]
    a = b = world["st0"]
    for it in range(8):
        a = o.optimization_multi(a, coupled=False)
        b = s.optimization(b)
        for u in range(U):
            assert np.max(np.abs(a[u]["spline"] - b[u]["spline"])) < 1e-6, (it, u)
            assert abs(a[u]["piece_time"] - b[u]["piece_time"]) < 1e-6
        assert abs(a[0]["gnorm"] - b[0]["gnorm"]) <= 1e-6 * max(1.0, a[0]["gnorm"])


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_sharded_equals_single(ranks):
    """needs `ranks` GPUs: robots sharded over the ranks with the native NCCL exchange (decoupled, coupled, unequal shares,
    overflow agreement, legacy callbacks) == one context bitwise, and tracks the oracle (tests/run_sharded_check.py)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + ranks), os.path.join(ROOT, "tests", "run_sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_CHECK ALL ranks=%d ok:True" % ranks in out.stdout, out.stdout[-3000:]


def test_coupled_iterations_against_golden():
    """mode 1 (Optimization3D_multi::optimization, one shared piece time) against tests/golden/coupled.npz, which travels
    without the reference"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coupled.npz"))
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    wps = sc["way_points"][:2] + sc["way_points"][4:6]
    s = api.Solver(4, uav_num=4, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in wps]
    for i in range(1, 7):
        sts = s.optimization(sts, coupled=True)
        for u, st in enumerate(sts):
            assert np.max(np.abs(st["spline"] - g["it%d_u%d_spline" % (i, u)])) < 1e-6, (i, u)
            assert np.max(np.abs(st["p_slack"] - g["it%d_u%d_p_slack" % (i, u)])) < 1e-6, (i, u)
        assert abs(sts[0]["piece_time"] - float(g["it%d_piece_time" % i])) < 1e-6
