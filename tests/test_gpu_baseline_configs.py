"""GPU parity at the sizes BASELINE.json names (not only the small scenes of the other test files): the CUDA path through
the C ABI against the compiled reference (oracle/_ref) on

  configs[1]  forest: 1 M points, 64 Bezier pieces (512 sub-segments, Newton system n = 574, 65-block cyclic reduction)
  configs[2]  cross: 8 UAVs, 50 k points, decoupled and coupled iterations
  configs[3]  circle: 64 UAVs with inter-robot planes, one GPU
  configs[4]  one member of the batched sweep with a ~0.9 M point cloud, run as a slot of a mode-2 batch
  P = 128     the banded fall-back of the Newton solve (k_solve, taken when the cyclic reduction does not fit shared memory)

Tolerances (BASELINE.json north_star): candidate pairs and planes bit-exact, energy / gradient 1e-9 relative, Newton
direction 1e-7, CCD step equal (never larger), trajectories 1e-6 over the tested iterations.
"""
import numpy as np
import pytest

from trajopt import api, scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def csr_sets_equal(a, b):
    (oa_, ia), (ob, ib) = a, b
    if not np.array_equal(oa_, ob):
        return False
    for r in range(len(oa_) - 1):
        if not np.array_equal(np.sort(ia[oa_[r]:oa_[r + 1]]), np.sort(ib[ob[r]:ob[r + 1]])):
            return False
    return True


def planes_sorted(off, c, d):
    out = []
    for r in range(len(off) - 1):
        blk = np.column_stack([c[off[r]:off[r + 1]], d[off[r]:off[r + 1]]])
        if len(blk):
            blk = blk[np.lexsort(blk.T[::-1])]
        out.append(blk)
    return out


# ---- configs[1]: forest ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def forest(oracle_ref):
    sc = scenes.forest()                       # 1 000 000 points, 64 pieces: the benchmarked configuration
    P = len(sc["way_points"][0]) - 1
    assert P == 64 and sc["V"].shape[0] == 1_000_000
    o = oracle_ref
    o.setup(oa.Params(P, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    st0 = scenes.initial_states(sc)[0]
    st = st0
    for _ in range(2):                         # a generic state (planes active, bound barriers active)
        st = o.optimization(st)
    yield dict(o=o, s=s, st0=st0, st=st, P=P)
    s.close()


@pytest.mark.parametrize("which", ["st0", "st"])
def test_forest_candidates_and_planes_bit_exact(forest, which):
    o, s, sp = forest["o"], forest["s"], forest[which]["spline"]
    ref = o.dcd_collision(sp, 0.2)
    got = s.dcd_collision(sp, 0.2)
    assert len(ref[1]) > 100_000
    assert csr_sets_equal(ref, got)
    ro, rc, rd = o.separate_plane(sp)
    go, gc, gd = s.separate_plane(sp)
    assert np.array_equal(ro, go) and len(rd) > 20_000
    for a, b in zip(planes_sorted(ro, rc, rd), planes_sorted(go, gc, gd)):
        assert np.array_equal(a, b)


def test_forest_energy_gradient_direction_step(forest):
    o, s, st, P = forest["o"], forest["s"], forest["st"], forest["P"]
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    e_ref = o.spline_energy(st, planes)
    assert np.isfinite(e_ref) and abs(s.spline_energy(st) - e_ref) <= 1e-9 * abs(e_ref)
    g_ref, h_ref = o.global_spline_gradient(st, planes)
    g, h = s.global_spline_gradient(st)
    assert rel(g, g_ref) < 1e-9 and rel(h, h_ref) < 1e-9
    # L1 at n = 3(T-4)+1 = 574: 65 blocks, 7 reduction levels (Optimization3D_admm.h:400-503)
    d_ref, td_ref, w_ref, gn_ref = o.descent_direction(st, planes)
    d, td, w, gn = s.descent_direction(st)
    assert rel(d, d_ref) < 1e-7
    assert abs(td - td_ref) <= 1e-7 * max(abs(td_ref), 1e-12)
    assert abs(w - w_ref) <= 1e-8 * abs(w_ref) and abs(gn - gn_ref) <= 1e-9 * abs(gn_ref)
    # swept broadphase + CCD ladder on the reference's own direction and on larger / random ones
    ref = o.ccd_collision(st["spline"], d_ref, 0.1)
    got = s.ccd_collision(st["spline"], d_ref, 0.1)
    assert csr_sets_equal(ref, got)
    rng = np.random.default_rng(7)
    dirs = [d_ref, 4.0 * d_ref]
    for _ in range(2):
        dd = np.zeros_like(d_ref); dd[2:-2] = rng.normal(size=(d_ref.shape[0] - 4, 3)) * 0.15
        dirs.append(np.asfortranarray(dd))
    seen = set()
    for dd in dirs:
        r = o.position_step(st["spline"], dd)
        g_ = s.position_step(st["spline"], dd)
        assert g_ <= r and g_ == r
        seen.add(r)
    assert len(seen) > 1


def test_forest_iterations_track_reference(forest):
    o, s = forest["o"], forest["s"]
    a = b = forest["st0"]
    for it in range(5):
        a = o.optimization(a)
        b = s.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6, it
        assert abs(a["piece_time"] - b["piece_time"]) < 1e-6
        assert abs(a["gnorm"] - b["gnorm"]) <= 1e-6 * max(1.0, abs(a["gnorm"]))
    for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
        assert np.max(np.abs(a[k] - b[k])) < 1e-6, k


# ---- configs[2]: cross, 8 UAVs, 50 k points ----------------------------------------------------------------------------
@pytest.mark.parametrize("coupled", [False, True])
def test_cross8_50k_iterations(oracle_ref, coupled):
    sc = scenes.cross(n_pts=50_000)
    U, P = sc["uav_num"], 8
    o = oracle_ref
    o.setup(oa.Params(P, uav_num=U, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, uav_num=U, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    a = b = scenes.initial_states(sc)
    for it in range(6):
        a = o.optimization_multi(a, coupled=coupled)
        b = s.optimization(b, coupled=coupled)
        for u in range(U):
            assert np.max(np.abs(a[u]["spline"] - b[u]["spline"])) < 1e-6, (it, u)
            assert abs(a[u]["piece_time"] - b[u]["piece_time"]) < 1e-6
        assert abs(a[0]["gnorm"] - b[0]["gnorm"]) <= 1e-6 * max(1.0, a[0]["gnorm"])
    assert s.counters()["planes"] > 0
    s.close()


# ---- configs[3]: circle, 64 UAVs, one GPU ------------------------------------------------------------------------------
def test_circle64_single_gpu(oracle_ref):
    sc = scenes.circle(n_uav=64, n_pts=20_000)
    U, P = 64, 8
    o = oracle_ref
    o.setup(oa.Params(P, uav_num=U, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, uav_num=U, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    st0 = scenes.initial_states(sc)
    # inter-robot planes of the initial state: same (row, plane) structure, c bit-exact, d out of the Newton loop on log()
    splines = [x["spline"] for x in st0]
    so, sc_, sd = o.separate_self(splines, cap=1 << 20)
    go, gc, gd = s.separate_planes(splines, with_self=True)
    n_self = 0
    ref_rows, got_rows = planes_sorted(so, sc_, sd), planes_sorted(go, gc, gd)
    for r in range(U * P * 8):
        # the circle scene has no obstacle candidates at the start: every plane of a row is an inter-robot plane
        assert len(ref_rows[r]) == len(got_rows[r]), r
        if len(ref_rows[r]):
            assert np.max(np.abs(ref_rows[r] - got_rows[r])) <= 1e-12
            n_self += len(ref_rows[r])
    assert n_self > 500
    a = b = st0
    for it in range(5):
        a = o.optimization_multi(a, coupled=False)
        b = s.optimization(b)
        for u in range(U):
            assert np.max(np.abs(a[u]["spline"] - b[u]["spline"])) < 1e-6, (it, u)
            assert abs(a[u]["piece_time"] - b[u]["piece_time"]) < 1e-6
    s.close()


# ---- configs[4]: a big member of the batched sweep as one slot of a mode-2 batch -----------------------------------------
def test_batch_member_with_large_cloud(oracle_ref):
    big = next(k for k in range(1024) if scenes.batch_member_meta(k)[0] > 800_000 and scenes.batch_member_meta(k)[1] < 0.2)
    small = [k for k in range(64) if scenes.batch_member_meta(k)[0] < 40_000][:2]
    ks = [small[0], big, small[1]]
    ms = [scenes.batch_member(k) for k in ks]
    P = 8
    s = api.Solver(P, uav_num=len(ms), ks=1e-8)
    s.init_pointclouds([m["V"] for m in ms])
    sts = [scenes.init_state(scenes.init_spline_single(m["way_points"][0])) for m in ms]
    s.states_upload(sts)
    iters = 3
    s.iterate(iters, mode=2)
    got = s.states_download(sts)
    o = oracle_ref
    for slot, m in enumerate(ms):
        o.setup(oa.Params(P, ks=1e-8))
        o.init_pointcloud(m["V"])
        a = sts[slot]
        if slot == 1:   # candidate sets of the big cloud, cloud-local ids
            ref = o.dcd_collision(a["spline"], 0.2, cap=1 << 22)
            assert len(ref[1]) > 500_000
        for _ in range(iters):
            a = o.optimization(a)
        assert np.max(np.abs(a["spline"] - got[slot]["spline"])) < 1e-6, slot
        assert abs(a["piece_time"] - got[slot]["piece_time"]) < 1e-6
        for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
            assert np.max(np.abs(a[k] - got[slot][k])) < 1e-6, (slot, k)
    s.close()


# ---- P = 128: banded fall-back of the Newton solve -------------------------------------------------------------------------
def test_long_trajectory_uses_banded_fallback(oracle_ref):
    """129 blocks x 288 doubles exceed the 220 KB of shared memory the cyclic reduction may use: solve_directions takes
    k_solve (sequential banded Cholesky, band in global memory).  n = 3(T-4)+1 = 1150."""
    sc = scenes.forest(n_pts=120_000, n_pieces=128)
    P = 128
    o = oracle_ref
    o.setup(oa.Params(P, ks=sc["ks"]))
    o.init_pointcloud(sc["V"])
    s = api.Solver(P, ks=sc["ks"])
    s.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    st = o.optimization(st)
    planes = o.separate_plane(st["spline"])
    s.set_planes(planes)
    d_ref, td_ref, w_ref, gn_ref = o.descent_direction(st, planes)
    d, td, w, gn = s.descent_direction(st)
    assert rel(d, d_ref) < 1e-7
    assert abs(td - td_ref) <= 1e-7 * max(abs(td_ref), 1e-12)
    assert abs(w - w_ref) <= 1e-8 * abs(w_ref) and abs(gn - gn_ref) <= 1e-9 * abs(gn_ref)
    a = b = scenes.initial_states(sc)[0]
    for it in range(3):
        a = o.optimization(a)
        b = s.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6, it
    s.close()
