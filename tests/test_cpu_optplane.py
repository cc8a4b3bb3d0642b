"""CPU-side tests of the persistent-plane mode ("optimal_plane": 1): both oracles and the host build of the product's
refinement header (csrc/optplane.cuh) against tests/golden/optplane.npz, which was produced by the unmodified reference
(tests/golden/make_golden.py optplane).

Tolerances: optimal_cd (segment vs point) is well conditioned: 1e-9 on (c, d).  self_optimal_cd shifts an indefinite 3x3
Hessian to a smallest eigenvalue of 1e-8 and may take thousands of tiny steps: the golden file marks the pairs whose
reference result survives a one-ulp perturbation of the input (`scd_stable`, 194 of 300); those are compared to 1e-6, all
pairs through the properties the algorithm guarantees (stopping criterion |grad| < 1e-2, feasibility, unit normal, barrier
energy not above the start and of the size of the reference's)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_api as oa
from trajopt import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)
OFFSET = MARGIN = 0.1


def D(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def go():
    return np.load(os.path.join(ROOT, "tests", "golden", "optplane.npz"))


def hulls_energy_grad(P0, P1, c, d):
    """Optimal_plane::self_barrier_energy / the gradient norm of self_barrier_grad (Optimal_plane.h:518-618), numpy"""
    c0 = np.array([c[1], -c[0], 0.0]); c0 /= np.linalg.norm(c0)
    e, g = 0.0, np.zeros(3)
    for Q, sg in ((P0, 1.0), (P1, -1.0)):
        for j in range(6):
            dist = sg * (Q[j] @ c) + sg * d - 0.5 * OFFSET
            if dist <= 0:
                return np.inf, np.inf
            if dist < MARGIN:
                e += -(dist - MARGIN) ** 2 * np.log(dist / MARGIN)
                e1 = -(2 * (dist - MARGIN) * np.log(dist / MARGIN) + (dist - MARGIN) ** 2 / dist)
                g += np.array([e1 * sg * (Q[j] @ c0), 0.0, sg * e1])
    return e, float(np.linalg.norm(g))


def check_self_pairs(go, refine):
    """refine(P0, P1, c0, d0) -> (c, d) for every golden pair"""
    n = len(go["scd_d1"])
    worst_stable = 0.0
    for i in range(n):
        P0, P1 = go["scd_P0"][i], go["scd_P1"][i]
        c, d = refine(P0, P1, go["scd_c0"][i], float(go["scd_d0"][i]))
        e, gn = hulls_energy_grad(P0, P1, c, d)
        e0, _ = hulls_energy_grad(P0, P1, go["scd_c0"][i], float(go["scd_d0"][i]))
        er, _ = hulls_energy_grad(P0, P1, go["scd_c1"][i], float(go["scd_d1"][i]))
        assert np.isfinite(e) and gn < 1e-2, (i, e, gn)              # the reference's only exit: |grad| < 1e-2
        assert abs(np.linalg.norm(c) - 1) < 1e-9
        assert e <= e0 * (1 + 1e-12) + 1e-15                          # Armijo steps never increase the barrier energy
        assert e <= 1.5 * er + 1e-3, (i, e, er)                       # ... and both stop in the same flat valley
        if go["scd_stable"][i]:
            worst_stable = max(worst_stable, np.abs(c - go["scd_c1"][i]).max(), abs(d - go["scd_d1"][i]))
    assert worst_stable < 1e-6, worst_stable


@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_optimal_cd_against_golden(go, kind):
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    o = oa.get(kind)
    o.setup(oa.Params(4, ks=1e-8))
    for i in range(len(go["cd_d1"])):
        c, d = o.optimal_cd(go["cd_P"][i], go["cd_q"][i], go["cd_c0"][i], float(go["cd_d0"][i]))
        tol = 0.0 if kind == "ref" else 1e-9
        assert np.abs(c - go["cd_c1"][i]).max() <= tol and abs(d - go["cd_d1"][i]) <= tol, i


@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_self_optimal_cd_against_golden(go, kind):
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    o = oa.get(kind)
    o.setup(oa.Params(4, uav_num=4, ks=1e-3))
    check_self_pairs(go, lambda P0, P1, c, d: o.self_optimal_cd(P0, P1, c, d))


@pytest.mark.parametrize("kind", ["ref", "port"])
def test_oracle_persistent_iterations_against_golden(go, kind):
    if kind not in oa.available():
        pytest.skip(kind + " oracle not built")
    sc = scenes.bridge(n_pts=4000, seed=21, n_pieces=4)
    o = oa.get(kind)
    o.setup(oa.Params(4, ks=sc["ks"], optimal_plane=1)); o.init_pointcloud(sc["V"]); o.reset_persistent_planes()
    st = scenes.initial_states(sc)[0]
    for i in range(1, 7):
        st = o.optimization(st)
        assert np.abs(st["spline"] - go["s_it%d_spline" % i]).max() < (1e-13 if kind == "ref" else 1e-9), i
        if i in (1, 3, 6):
            tr, ids, c, d = o.live_planes()
            assert np.array_equal(tr, go["s_it%d_live_tr" % i]) and np.array_equal(ids, go["s_it%d_live_id" % i])
            assert np.abs(c - go["s_it%d_live_c" % i]).max() < 1e-9 and np.abs(d - go["s_it%d_live_d" % i]).max() < 1e-9
    assert len(go["s_it6_live_d"]) > 300
    o.setup(oa.Params(4, ks=sc["ks"]))      # leave the shared oracle in the default mode


def test_port_persistent_multi_against_golden(go):
    """decoupled 4-UAV iterations with persistent inter-robot planes (Optimization3D_multi.h:271-339)"""
    if "port" not in oa.available():
        pytest.skip("port oracle not built")
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    wps = sc["way_points"][:2] + sc["way_points"][4:6]
    o = oa.get("port")
    o.setup(oa.Params(4, uav_num=4, ks=sc["ks"], optimal_plane=1)); o.init_pointcloud(sc["V"]); o.reset_persistent_planes()
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in wps]
    for i in range(1, 5):
        sts = o.optimization_multi(sts, coupled=False)
        for u, s in enumerate(sts):
            assert np.abs(s["spline"] - go["m_it%d_u%d_spline" % (i, u)]).max() < 1e-6, (i, u)
    o.setup(oa.Params(4, ks=sc["ks"]))


def test_hostsim_optimal_cd_against_golden(hostsim, go):
    """the product's csrc/optplane.cuh compiled by g++: same refinement as the reference on the golden pairs"""
    for i in range(len(go["cd_d1"])):
        P = np.asfortranarray(go["cd_P"][i]); q = np.ascontiguousarray(go["cd_q"][i])
        c = go["cd_c0"][i].copy(); d = C.c_double(float(go["cd_d0"][i]))
        rc = hostsim.hs_optimal_cd(D(P), D(q), C.c_double(OFFSET), C.c_double(MARGIN), D(c), C.byref(d))
        assert rc == 0
        assert np.abs(c - go["cd_c1"][i]).max() < 1e-12 and abs(d.value - go["cd_d1"][i]) < 1e-12, i


def test_hostsim_self_optimal_cd_against_golden(hostsim, go):
    def refine(P0, P1, c0, d0):
        A = np.asfortranarray(P0); B = np.asfortranarray(P1)
        c = c0.copy(); d = C.c_double(d0)
        assert hostsim.hs_self_optimal_cd(D(A), D(B), C.c_double(OFFSET), C.c_double(MARGIN), D(c), C.byref(d)) == 0
        return c, d.value
    check_self_pairs(go, refine)
