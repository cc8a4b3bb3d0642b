"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): candidate pairs bit-exact; planes bit-exact (same arithmetic, no FMA);
energy / gradient 1e-9 relative; CCD step never larger than the reference's (here: equal); trajectories 1e-6.
"""
import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu

P = 8


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def world(oracle_any):
    sc = scenes.bridge(n_pts=30000, seed=11)
    o = oracle_any
    o.setup(oa.Params(P, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    st0 = scenes.initial_states(sc)[0]
    # a generic (non-straight) state: three oracle iterations
    st = st0
    for _ in range(3):
        st = o.optimization(st)
    return dict(sc=sc, o=o, s=s, st0=st0, st=st)


def csr_sets_equal(a, b):
    (oa_, ia), (ob, ib) = a, b
    if not np.array_equal(oa_, ob):
        return False
    for r in range(len(oa_) - 1):
        if not np.array_equal(np.sort(ia[oa_[r]:oa_[r + 1]]), np.sort(ib[ob[r]:ob[r + 1]])):
            return False
    return True


def test_tables_bit_exact(world):
    to, ts = world["o"].tables(), world["s"].tables()
    for k in to:
        assert np.array_equal(to[k], ts[k]), k


@pytest.mark.parametrize("which", ["st0", "st"])
def test_broadphase_dcd_bit_exact(world, which):
    sp = world[which]["spline"]
    ref = world["o"].dcd_collision(sp, 0.2)
    got = world["s"].dcd_collision(sp, 0.2)
    assert len(ref[1]) > 1000
    assert csr_sets_equal(ref, got)


def test_broadphase_ccd_bit_exact(world):
    st = world["st"]
    planes = world["o"].separate_plane(st["spline"])
    direction, *_ = world["o"].descent_direction(st, planes)
    ref = world["o"].ccd_collision(st["spline"], direction, 0.1)
    got = world["s"].ccd_collision(st["spline"], direction, 0.1)
    assert csr_sets_equal(ref, got)


def sort_planes(off, c, d):
    out = []
    for r in range(len(off) - 1):
        blk = np.column_stack([c[off[r]:off[r + 1]], d[off[r]:off[r + 1]]])
        if len(blk):
            blk = blk[np.lexsort(blk.T[::-1])]
        out.append(blk)
    return out


@pytest.mark.parametrize("which", ["st0", "st"])
def test_separate_plane_bit_exact(world, which):
    sp = world[which]["spline"]
    ro, rc, rd = world["o"].separate_plane(sp)
    go, gc, gd = world["s"].separate_plane(sp)
    assert np.array_equal(ro, go)
    assert len(rd) > 500
    for a, b in zip(sort_planes(ro, rc, rd), sort_planes(go, gc, gd)):
        assert np.array_equal(a, b)


def test_energies(world):
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    e_ref = o.spline_energy(st, planes)
    assert np.isfinite(e_ref)
    assert abs(s.spline_energy(st) - e_ref) <= 1e-9 * abs(e_ref)
    b_ref = o.plane_barrier_energy(st["spline"], planes)
    assert abs(s.plane_barrier_energy(st["spline"]) - b_ref) <= 1e-9 * abs(b_ref)
    for t in (st["piece_time"], 0.9):
        bo = o.bound_energy(st["spline"], t)
        bs = s.bound_energy(st["spline"], t)
        assert (np.isinf(bo) and np.isinf(bs)) or abs(bs - bo) <= 1e-9 * max(abs(bo), 1e-12)


def test_energy_infeasible_is_inf(world):
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    bad = dict(st); sp = st["spline"].copy(order="F"); sp[5:20, 2] -= 0.3; bad["spline"] = sp
    assert np.isinf(o.spline_energy(bad, planes))
    assert np.isinf(s.spline_energy(bad))


def test_piece_blocks(world):
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    g, h = s.piece_blocks(st, project_psd=False)
    for sp in range(P):
        g0, h0 = o.local_spline_gradient(st, planes, sp)
        assert rel(g[sp], g0) < 1e-9
        assert rel(h[sp], h0) < 1e-9


def test_global_gradient_and_direction(world):
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    g_ref, h_ref = o.global_spline_gradient(st, planes)
    g, h = s.global_spline_gradient(st)
    assert rel(g, g_ref) < 1e-9 and rel(h, h_ref) < 1e-9
    d_ref, td_ref, w_ref, gn_ref = o.descent_direction(st, planes)
    d, td, w, gn = s.descent_direction(st)
    assert rel(d, d_ref) < 1e-7
    assert abs(td - td_ref) <= 1e-7 * max(abs(td_ref), 1e-12)
    assert abs(w - w_ref) <= 1e-8 * abs(w_ref)
    assert abs(gn - gn_ref) <= 1e-9 * abs(gn_ref)


def test_psd_projection_of_indefinite_blocks(world):
    """Gradient_admm.h:40-53: a piece block that fails Eigen's LLT is shifted by (-lambda_min + 0.01) I.  Short piece times
    switch the velocity / acceleration barriers on, whose time coupling makes blocks indefinite; the kernel decides by
    lambda_min first and must reproduce the reference's shifted blocks."""
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    shifted = 0
    for t in np.linspace(2.10, 2.20, 21):          # just above the shortest feasible piece time of this state (~2.12)
        st2 = dict(st, piece_time=float(t))
        if not np.isfinite(o.spline_energy(st2, planes)):
            continue
        g_ref, h_ref = o.global_spline_gradient(st2, planes)
        g, h = s.global_spline_gradient(st2)
        assert rel(g, g_ref) < 1e-9 and rel(h, h_ref) < 1e-9, t
        _, h_raw = s.piece_blocks(st2, project_psd=False)
        _, h_psd = s.piece_blocks(st2, project_psd=True)
        n = sum(1 for sp in range(P) if not np.array_equal(h_raw[sp], h_psd[sp]))
        for sp in range(P):
            if not np.array_equal(h_raw[sp], h_psd[sp]):      # shifted: exactly a multiple of the identity, lambda_min + 0.01 > 0
                dlt = h_psd[sp] - h_raw[sp]
                assert np.allclose(dlt, dlt[0, 0] * np.eye(19), rtol=0, atol=1e-12 * abs(dlt[0, 0])) and dlt[0, 0] > 0.01
                assert abs(np.linalg.eigvalsh(h_psd[sp])[0] - 0.01) < 1e-9 * max(1.0, np.abs(h_raw[sp]).max())
        shifted += n
        d_ref, td_ref, w_ref, gn_ref = o.descent_direction(st2, planes)
        d, td, w, gn = s.descent_direction(st2)
        assert rel(d, d_ref) < 1e-7, t
    assert shifted > 0


def test_position_step_equals_reference(world):
    st, o, s = world["st"], world["o"], world["s"]
    planes = o.separate_plane(st["spline"])
    direction, *_ = o.descent_direction(st, planes)
    rng = np.random.default_rng(5)
    dirs = [direction, 3.0 * direction]
    for _ in range(3):
        dd = np.zeros_like(direction); dd[2:-2] = rng.normal(size=(direction.shape[0] - 4, 3)) * 0.2
        dirs.append(np.asfortranarray(dd))
    seen = set()
    for dd in dirs:
        ref = o.position_step(st["spline"], dd)
        got = s.position_step(st["spline"], dd)
        assert got <= ref
        assert got == ref
        seen.add(ref)
    assert len(seen) > 1


def test_update_slack_lambda(world):
    st, o, s = world["st"], world["o"], world["s"]
    a = o.update_slack_lambda(st)
    b = s.update_slack_lambda(st)
    for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
        assert rel(b[k], a[k]) < 1e-9, k


def test_iterations_track_reference(world):
    o, s = world["o"], world["s"]
    a = b = world["st0"]
    for it in range(10):
        a = o.optimization(a)
        b = s.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6, it
        assert abs(a["piece_time"] - b["piece_time"]) < 1e-6
        assert abs(a["gnorm"] - b["gnorm"]) <= 1e-6 * max(1.0, abs(a["gnorm"]))
    for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
        assert np.max(np.abs(a[k] - b[k])) < 1e-6, k


def test_resident_iterations_match_host_roundtrip(world):
    s = world["s"]
    st = world["st0"]
    s.states_upload([st])
    s.iterate(3)
    res = s.states_download([st])[0]
    b = st
    for _ in range(3):
        b = s.optimization(b)
    assert np.array_equal(res["spline"], b["spline"])
    assert res["piece_time"] == b["piece_time"]
