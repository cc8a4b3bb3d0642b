"""Drop-in tests of the C++ host layer (traj-opt-admm_b200/host): the reference's own callers, compiled UNCHANGED against
the shadow headers and linked with the CUDA library, against the same callers compiled against the reference.

  * oracle/ref_shim.cpp (flat-C driver over Energy_admm / Gradient_admm / Step / Separate / Optimal_plane / CCD / BVH /
    Optimization3D_*) exists twice: oracle/_ref/libtrajopt_ref.so (reference headers, CPU) and
    traj-opt-admm_b200/host/build/libtrajopt_hostshim.so (shadow headers -> C ABI -> GPU).  Same ctypes front-end for both.
  * Main/admmPathPlanning3D.cpp and Main/multiPathPlanning3D.cpp exist twice as executables; both run on the same scene
    files (reference formats) and must stop at the same iteration with the same trajectory.

The binaries are built where /root/reference exists (__graft_entry__.build()) and travel to the GPU box prebuilt.
Tolerances: discrete outputs equal, planes bit-exact, energies/gradients 1e-9 relative, trajectories 1e-6.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from trajopt import scenes
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "traj-opt-admm_b200", "host", "build")
HOSTSHIM = os.path.join(HOST, "libtrajopt_hostshim.so")
P = 8


class HostDropIn(oa._Base):
    """oracle/ref_shim.cpp compiled against the shadow headers: every ref_* call lands in the CUDA library"""
    prefix = "ref_"
    kind = "b200-host"

    def __init__(self):
        super().__init__(HOSTSHIM)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def sorted_rows(off, c, d):
    out = []
    for r in range(len(off) - 1):
        blk = np.column_stack([c[off[r]:off[r + 1]], d[off[r]:off[r + 1]]])
        if len(blk):
            blk = blk[np.lexsort(blk.T[::-1])]
        out.append(blk)
    return out


@pytest.fixture(scope="module")
def pair(oracle_ref):
    if not os.path.exists(HOSTSHIM):
        pytest.skip("host drop-in not built (needs /root/reference at build time)")
    sc = scenes.bridge(n_pts=20000, seed=21)
    ref, dev = oracle_ref, HostDropIn()
    for o in (ref, dev):
        o.setup(oa.Params(P, ks=sc["ks"]))
        o.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    for _ in range(3):
        st = ref.optimization(st)
    return dict(ref=ref, dev=dev, st=st, st0=scenes.initial_states(sc)[0], sc=sc)


def test_bvh_queries(pair):
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    a, b = ref.dcd_collision(st["spline"], 0.2), dev.dcd_collision(st["spline"], 0.2)
    assert np.array_equal(a[0], b[0]) and len(a[1]) > 500
    for r in range(len(a[0]) - 1):
        assert np.array_equal(np.sort(a[1][a[0][r]:a[0][r + 1]]), b[1][b[0][r]:b[0][r + 1]])
    planes = ref.separate_plane(st["spline"])
    direction = ref.descent_direction(st, planes)[0]
    a, b = ref.ccd_collision(st["spline"], direction, 0.1), dev.ccd_collision(st["spline"], direction, 0.1)
    assert np.array_equal(a[0], b[0])
    for r in range(len(a[0]) - 1):
        assert np.array_equal(np.sort(a[1][a[0][r]:a[0][r + 1]]), b[1][b[0][r]:b[0][r + 1]])


def test_ccd_and_plane_primitives(pair):
    ref, dev, st, sc = pair["ref"], pair["dev"], pair["st"], pair["sc"]
    rng = np.random.default_rng(3)
    off, ids = ref.dcd_collision(st["spline"], 0.2)
    n_plane = 0
    for tr in rng.choice(P * 8, size=12, replace=False):
        Pm = ref.segment_points(st["spline"], int(tr))
        assert np.array_equal(Pm, dev.segment_points(st["spline"], int(tr)))
        for pid in ids[off[tr]:off[tr + 1]][:25]:
            q = sc["V"][pid:pid + 1]
            assert ref.kdop_dcd(Pm, q, 0.2) == dev.kdop_dcd(Pm, q, 0.2)
            assert np.array_equal(ref.gjk(Pm, q), dev.gjk(Pm, q))
            ra, rb = ref.opengjk(Pm, q, 0.2), dev.opengjk(Pm, q, 0.2)
            assert ra[0] == rb[0]
            if ra[0]:
                n_plane += 1
                assert np.array_equal(ra[1], rb[1]) and ra[2] == rb[2]
            D = rng.normal(size=(6, 3)) * 0.3
            assert ref.kdop_ccd(Pm, D, q, 0.1, 0.0, 0.7) == dev.kdop_ccd(Pm, D, q, 0.1, 0.0, 0.7)
            assert ref.gjk_ccd(Pm, D, q, 0.1, 0.0, 0.7) == dev.gjk_ccd(Pm, D, q, 0.1, 0.0, 0.7)
    assert n_plane >= 3
    # segment vs segment
    for _ in range(10):
        P0 = rng.normal(size=(6, 3)) * 0.2; P1 = rng.normal(size=(6, 3)) * 0.2 + np.array([0.5, 0.1, 0.0])
        D0 = rng.normal(size=(6, 3)) * 0.3; D1 = rng.normal(size=(6, 3)) * 0.3
        assert ref.self_kdop_dcd(P0, P1, 0.3) == dev.self_kdop_dcd(P0, P1, 0.3)
        assert ref.self_kdop_ccd(P0, D0, P1, D1, 0.1, 0, 0.8, 0, 0.64) == dev.self_kdop_ccd(P0, D0, P1, D1, 0.1, 0, 0.8, 0, 0.64)
        assert ref.self_gjk_ccd(P0, D0, P1, D1, 0.1, 0, 0.8, 0, 0.64) == dev.self_gjk_ccd(P0, D0, P1, D1, 0.1, 0, 0.8, 0, 0.64)
        ra, rb = ref.selfgjk(P0, P1, 1.0), dev.selfgjk(P0, P1, 1.0)
        assert ra[0] == rb[0]
        if ra[0]:
            assert np.array_equal(ra[1], rb[1]) and ra[2] == rb[2]
            da, db = ref.optimal_d(P0, P1, ra[1], ra[2]), dev.optimal_d(P0, P1, ra[1], ra[2])
            # random hulls may sit closer than offset: the reference's Newton loop then runs into log(<0) = NaN, and so must we
            assert (np.isnan(da) and np.isnan(db)) or abs(da - db) <= 1e-12


def test_planes_energy_gradient_direction(pair):
    ref, dev, st = pair["ref"], pair["dev"], pair["st"]
    pr, pd = ref.separate_plane(st["spline"]), dev.separate_plane(st["spline"])
    assert np.array_equal(pr[0], pd[0]) and len(pr[2]) > 300
    for a, b in zip(sorted_rows(*pr), sorted_rows(*pd)):
        assert np.array_equal(a, b)
    e = ref.spline_energy(st, pr)
    assert abs(dev.spline_energy(st, pr) - e) <= 1e-9 * abs(e)
    e = ref.plane_barrier_energy(st["spline"], pr)
    assert abs(dev.plane_barrier_energy(st["spline"], pr) - e) <= 1e-9 * abs(e)
    e = ref.bound_energy(st["spline"], st["piece_time"])
    assert abs(dev.bound_energy(st["spline"], st["piece_time"]) - e) <= 1e-9 * max(abs(e), 1e-12)
    for sp in (0, 3, P - 1):
        g0, h0 = ref.local_spline_gradient(st, pr, sp)
        g1, h1 = dev.local_spline_gradient(st, pr, sp)
        assert rel(g1, g0) < 1e-9 and rel(h1, h0) < 1e-9
    g0, h0 = ref.global_spline_gradient(st, pr)
    g1, h1 = dev.global_spline_gradient(st, pr)
    assert rel(g1, g0) < 1e-9 and rel(h1, h0) < 1e-9
    d0, d1 = ref.descent_direction(st, pr), dev.descent_direction(st, pr)
    assert rel(d1[0], d0[0]) < 1e-7 and abs(d1[1] - d0[1]) <= 1e-7 * max(abs(d0[1]), 1e-12)
    assert abs(d1[2] - d0[2]) <= 1e-8 * abs(d0[2]) and abs(d1[3] - d0[3]) <= 1e-9 * abs(d0[3])
    assert ref.position_step(st["spline"], 3.0 * d0[0]) == dev.position_step(st["spline"], 3.0 * d0[0])
    a, b = ref.update_slack_lambda(st), dev.update_slack_lambda(st)
    for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
        assert rel(b[k], a[k]) < 1e-9, k


def test_single_uav_iterations(pair):
    ref, dev = pair["ref"], pair["dev"]
    a = b = pair["st0"]
    for it in range(8):
        a, b = ref.optimization(a), dev.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6, it
        assert abs(a["gnorm"] - b["gnorm"]) <= 1e-6 * max(1.0, abs(a["gnorm"]))


def test_persistent_plane_mode_through_the_shadow_headers(oracle_ref):
    """is_optimal_plane = 1: the reference's caller code (Optimization3D_admm::optimization, separate_plane and the
    Optimal_plane statics) against the shadow headers, live planes kept by the device context"""
    if not os.path.exists(HOSTSHIM):
        pytest.skip("host drop-in not built")
    sc = scenes.bridge(n_pts=8000, seed=3)
    ref, dev = oracle_ref, HostDropIn()
    for o in (ref, dev):
        o.setup(oa.Params(P, ks=sc["ks"], optimal_plane=1))
        o.init_pointcloud(sc["V"])
        o.reset_persistent_planes()
    a = b = scenes.initial_states(sc)[0]
    for it in range(5):
        a, b = ref.optimization(a), dev.optimization(b)
        assert np.max(np.abs(a["spline"] - b["spline"])) < 1e-6, it
        ta, ia, ca, da = ref.live_planes()
        tb, ib, cb, db = dev.live_planes()
        k = np.lexsort((ib, tb))
        assert np.array_equal(ta, tb[k]) and np.array_equal(ia, ib[k]), it
        assert np.max(np.abs(ca - cb[k])) < 1e-8 and np.max(np.abs(da - db[k])) < 1e-8
    assert len(ta) > 1500
    # the per-pair statics: Optimal_plane::optimal_cd / self_optimal_cd through the shadow header
    Pm = ref.segment_points(a["spline"], 20)
    q = sc["V"][int(ia[np.searchsorted(ta, 20)])] if (ta == 20).any() else sc["V"][int(ia[0])]
    ok, c0, d0 = ref.opengjk(Pm, q.reshape(1, 3), 10.0)
    c1, d1 = ref.optimal_cd(Pm, q, c0, d0)
    c2, d2 = dev.optimal_cd(Pm, q, c0, d0)
    assert np.max(np.abs(c1 - c2)) < 1e-9 and abs(d1 - d2) < 1e-9
    for o in (ref, dev):
        o.setup(oa.Params(P, ks=sc["ks"]))      # back to the default mode for the other tests


@pytest.mark.parametrize("coupled", [False, True])
def test_multi_uav_iterations(oracle_ref, coupled):
    if not os.path.exists(HOSTSHIM):
        pytest.skip("host drop-in not built")
    sc = scenes.cross(n_pts=10000, seed=3)
    U = sc["uav_num"]
    ref, dev = oracle_ref, HostDropIn()
    for o in (ref, dev):
        o.setup(oa.Params(P, uav_num=U, ks=sc["ks"]))
        o.init_pointcloud(sc["V"])
    a = b = scenes.initial_states(sc)
    splines = [x["spline"] for x in a]
    sr, sd = ref.separate_self(splines), dev.separate_self(splines)
    assert np.array_equal(sr[0], sd[0]) and len(sr[2]) > 0
    for x, y in zip(sorted_rows(*sr), sorted_rows(*sd)):
        assert len(x) == len(y) and (len(x) == 0 or np.max(np.abs(x - y)) <= 1e-12)
    for it in range(6):
        a, b = ref.optimization_multi(a, coupled=coupled), dev.optimization_multi(b, coupled=coupled)
        for u in range(U):
            assert np.max(np.abs(a[u]["spline"] - b[u]["spline"])) < 1e-6, (it, u)
            assert abs(a[u]["piece_time"] - b[u]["piece_time"]) < 1e-6
        assert abs(a[0]["gnorm"] - b[0]["gnorm"]) <= 1e-6 * max(1.0, a[0]["gnorm"])


def _run_main(exe, root, name):
    out = subprocess.run([exe, name], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    vals = {}
    for key in ("ccd time", "ccd len"):
        m = re.findall(r"^%s:([-+0-9.eE]+)" % key, out.stdout, flags=re.M)
        vals[key] = [float(x) for x in m]
    return vals


def _result_iter(path):
    m = re.search(r"iter: (\d+)", open(path).read())
    return int(m.group(1)) if m else None


def test_admm_executable_links_unchanged_and_matches(tmp_path):
    """Main/admmPathPlanning3D.cpp, unchanged: reference build vs shadow-header build on the same files"""
    exe_ref = os.path.join(ROOT, "oracle", "_ref", "admmPathPlanning3D_ref")
    exe_dev = os.path.join(HOST, "admmPathPlanning3D")
    if not (os.path.exists(exe_ref) and os.path.exists(exe_dev)):
        pytest.skip("executables not built (need /root/reference at build time)")
    sc = scenes.bridge(n_pts=6000, seed=5)
    out = {}
    for tag, exe in (("ref", exe_ref), ("dev", exe_dev)):
        root = str(tmp_path / tag)
        scenes.write_reference_files(sc, root, "bridge.obj", {"stop": 1e-2, "exit": 1})
        out[tag] = _run_main(exe, root, "bridge.obj")
        out[tag]["iter"] = _result_iter(os.path.join(root, "result", "bridge.obj_result_file_admm.txt"))
    assert out["ref"]["iter"] is not None and out["ref"]["iter"] > 10
    # Round-off differences (1e-15 after one iteration) are amplified ~10x per iteration while the planes still switch
    # (profiles/r01_divergence_bridge_single.txt: up to 1e-3 mid-run) and contract again near the optimum: both builds stop
    # at the same iteration here, a slack of 3 keeps the test from being a coin flip on the stopping threshold.
    assert abs(out["dev"]["iter"] - out["ref"]["iter"]) <= 3
    for key in ("ccd time", "ccd len"):
        assert len(out["ref"][key]) == 1
        # the two builds print with different stream precision (the reference sets cout.precision(10) inside its line search)
        assert abs(out["dev"][key][0] - out["ref"][key][0]) <= 2e-5 * abs(out["ref"][key][0]), key


@pytest.mark.parametrize("decouple", [1, 0])
def test_multi_executable_links_unchanged_and_matches(tmp_path, decouple):
    exe_ref = os.path.join(ROOT, "oracle", "_ref", "multiPathPlanning3D_ref")
    exe_dev = os.path.join(HOST, "multiPathPlanning3D")
    if not (os.path.exists(exe_ref) and os.path.exists(exe_dev)):
        pytest.skip("executables not built (need /root/reference at build time)")
    sc = scenes.cross(n_pts=4000, seed=3)
    out = {}
    for tag, exe in (("ref", exe_ref), ("dev", exe_dev)):
        root = str(tmp_path / (tag + str(decouple)))
        scenes.write_reference_files(sc, root, "cross.obj", {"stop": 1.0 if decouple else 0.5, "exit": 1, "decouple": decouple})
        out[tag] = _run_main(exe, root, "cross.obj")
        out[tag]["iter"] = _result_iter(os.path.join(root, "result", "cross.obj_result_file_multi.txt"))
    assert out["ref"]["iter"] is not None and out["ref"]["iter"] > 2
    # The 8-robot iteration is chaotic in the numerical sense: a 1e-15 difference grows ~10x per iteration until the two
    # runs follow different (equally valid) sequences of plane switches (profiles/r01_divergence_cross_*.txt; the first 6
    # iterations are checked to 1e-6 in test_multi_uav_iterations).  The executables must therefore agree on the outcome,
    # not on the path: both converge, in a comparable number of iterations, to trajectories of the same length and duration.
    assert out["dev"]["iter"] is not None
    assert 0.5 * out["ref"]["iter"] <= out["dev"]["iter"] <= 2.0 * out["ref"]["iter"]
    assert len(out["ref"]["ccd len"]) == sc["uav_num"] and len(out["dev"]["ccd len"]) == sc["uav_num"]
    for key in ("ccd time", "ccd len"):
        assert np.allclose(out["dev"][key], out["ref"][key], rtol=2e-2, atol=0), key


@pytest.mark.parametrize("decouple", [1, 0])
def test_multi_executable_on_two_gpus_equals_one_gpu(tmp_path, decouple):
    """Main/multiPathPlanning3D.cpp, unchanged, with TRAJOPT_B200_GPUS=2: the session shards the robots over two GPUs (one
    context + one host thread per GPU, NCCL exchange inside the library) and must reproduce the one-GPU run exactly:
    same stopping iteration, same printed trajectory lengths / durations"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(HOST, "multiPathPlanning3D")
    if not os.path.exists(exe):
        pytest.skip("executable not built (needs /root/reference at build time)")
    sc = scenes.cross(n_pts=4000, seed=3)
    out = {}
    for gpus in (1, 2):
        root = str(tmp_path / ("g%d_%d" % (gpus, decouple)))
        scenes.write_reference_files(sc, root, "cross.obj", {"stop": 1.0 if decouple else 0.5, "exit": 1, "decouple": decouple})
        env = dict(os.environ, TRAJOPT_B200_GPUS=str(gpus))
        r = subprocess.run([exe, "cross.obj"], cwd=root, capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        vals = {key: re.findall(r"^%s:([-+0-9.eE]+)" % key, r.stdout, flags=re.M) for key in ("ccd time", "ccd len")}
        out[gpus] = (vals, _result_iter(os.path.join(root, "result", "cross.obj_result_file_multi.txt")))
    assert out[1][1] is not None and out[1][1] > 2
    assert out[1] == out[2]
