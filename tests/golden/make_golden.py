"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref (needs /root/reference at
build time; run in the build container: `python tests/golden/make_golden.py`).  The vectors pin both the plain-C
port (oracle/port) and the CUDA path on machines where the reference is not available.

single.npz : 1 UAV, bridge-shaped cloud (4000 pts, seed 21), 4 pieces
multi.npz  : 4 UAVs crossing, floor/ceiling cloud (3000 pts, seed 23), 4 pieces
coupled.npz  : the 4-UAV scene through Optimization3D_multi::optimization (coupled: one shared piece time)
optplane.npz : the same two scenes run with "optimal_plane": 1 (persistent planes refined by Optimal_plane::optimal_cd /
             self_optimal_cd), plus per-pair refinement vectors
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)
from trajopt import scenes  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def single():
    sc = scenes.bridge(n_pts=4000, seed=21, n_pieces=4)
    P = 4
    o = oa.RefOracle(); o.setup(oa.Params(P, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    out = dict(V=sc["V"], way_points=sc["way_points"][0], ks=sc["ks"])
    for k, v in o.tables().items():
        out["tab_" + k] = v
    st = scenes.initial_states(sc)[0]
    states = [st]
    for _ in range(6):
        st = o.optimization(st)
        states.append(st)
    for i, s in enumerate(states):
        for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
            out["it%d_%s" % (i, k)] = s[k]
        out["it%d_piece_time" % i] = s["piece_time"]
        if i:
            out["it%d_gnorm" % i] = s["gnorm"]
    st = states[2]
    sp = st["spline"]
    off, ids = o.dcd_collision(sp, 0.2)
    out["dcd_off"] = off; out["dcd_ids"] = np.concatenate([np.sort(ids[off[r]:off[r + 1]]) for r in range(len(off) - 1)]).astype(np.uint32)
    po, pc, pd = o.separate_plane(sp)
    out["pl_off"] = po; out["pl_c"] = pc; out["pl_d"] = pd
    out["e_spline"] = o.spline_energy(st, (po, pc, pd))
    out["e_barrier"] = o.plane_barrier_energy(sp, (po, pc, pd))
    out["e_bound"] = o.bound_energy(sp, st["piece_time"])
    g, h = o.global_spline_gradient(st, (po, pc, pd))
    out["grad"] = g; out["hess"] = h
    gl, hl = zip(*[o.local_spline_gradient(st, (po, pc, pd), i) for i in range(P)])
    out["local_g"] = np.array(gl); out["local_h"] = np.array(hl)
    d, td, w, gn = o.descent_direction(st, (po, pc, pd))
    out["dir"] = d; out["tdir"] = td; out["wolfe"] = w; out["gnorm"] = gn
    co, ci = o.ccd_collision(sp, d, 0.1)
    out["ccd_off"] = co; out["ccd_ids"] = np.concatenate([np.sort(ci[co[r]:co[r + 1]]) for r in range(len(co) - 1)] + [np.zeros(0, np.uint32)]).astype(np.uint32)
    rng = np.random.default_rng(7)
    dirs, steps = [d, 4.0 * d], []
    for _ in range(4):
        dd = np.zeros_like(d); dd[2:-2] = rng.normal(size=(d.shape[0] - 4, 3)) * 0.25
        dirs.append(np.asfortranarray(dd))
    for dd in dirs:
        steps.append(o.position_step(sp, dd))
    out["step_dirs"] = np.array(dirs); out["steps"] = np.array(steps)
    sl = o.update_slack_lambda(st)
    for k in ("p_slack", "t_slack", "p_lambda", "t_lambda"):
        out["slack_" + k] = sl[k]
    # per-pair primitives on real candidates
    Pm = np.array([o.segment_points(sp, tr) for tr in range(P * 8)])
    pairs = [(tr, pid) for tr in range(P * 8) for pid in ids[off[tr]:off[tr + 1]]]
    pairs = [pairs[i] for i in rng.choice(len(pairs), size=min(600, len(pairs)), replace=False)]
    out["pp_P"] = np.array([Pm[tr] for tr, _ in pairs]); out["pp_q"] = np.array([sc["V"][pid] for _, pid in pairs])
    out["pp_kdop"] = np.array([o.kdop_dcd(Pm[tr], sc["V"][pid].reshape(1, 3), 0.2) for tr, pid in pairs])
    out["pp_gjk"] = np.array([o.gjk(Pm[tr], sc["V"][pid].reshape(1, 3)) for tr, pid in pairs])
    res = [o.opengjk(Pm[tr], sc["V"][pid].reshape(1, 3), 0.2) for tr, pid in pairs]
    out["pp_ok"] = np.array([r[0] for r in res]); out["pp_c"] = np.array([r[1] for r in res]); out["pp_d"] = np.array([r[2] for r in res])
    np.savez_compressed(os.path.join(HERE, "single.npz"), **out)
    print("single.npz: candidates", len(ids), "planes", len(pd), "steps", steps)


def multi():
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    sc["way_points"] = sc["way_points"][:2] + sc["way_points"][4:6]   # 2 + 2 crossing lanes
    U, P = 4, 4
    o = oa.RefOracle(); o.setup(oa.Params(P, uav_num=U, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    out = dict(V=sc["V"], way_points=np.array(sc["way_points"]), ks=sc["ks"])
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in sc["way_points"]]
    hist = [sts]
    for _ in range(5):
        sts = o.optimization_multi(sts, coupled=False)
        hist.append(sts)
    for i, ss in enumerate(hist):
        for u, s in enumerate(ss):
            for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
                out["it%d_u%d_%s" % (i, u, k)] = s[k]
            out["it%d_u%d_piece_time" % (i, u)] = s["piece_time"]
        if i:
            out["it%d_gnorm" % i] = ss[0]["gnorm"]
    ss = hist[2]
    splines = [s["spline"] for s in ss]
    so, sc_, sd = o.separate_self(splines)
    out["self_off"] = so; out["self_c"] = sc_; out["self_d"] = sd
    # hull-hull primitives per time slot
    P0s, P1s, oks, cs, ds, dref = [], [], [], [], [], []
    for tr in range(P * 8):
        Pl = [o.segment_points(s, tr) for s in splines]
        for a in range(U):
            for b in range(a + 1, U):
                ok, c, d = o.selfgjk(Pl[a], Pl[b], 0.3)
                P0s.append(Pl[a]); P1s.append(Pl[b]); oks.append(ok); cs.append(c); ds.append(d)
                dref.append(o.optimal_d(Pl[a], Pl[b], c, d) if ok else 0.0)
    out["hh_P0"] = np.array(P0s); out["hh_P1"] = np.array(P1s); out["hh_ok"] = np.array(oks); out["hh_c"] = np.array(cs)
    out["hh_d"] = np.array(ds); out["hh_dref"] = np.array(dref)
    # self CCD step with synthetic directions
    rng = np.random.default_rng(9)
    dirs = []
    for s in splines:
        dd = np.zeros_like(s); dd[2:-2] = rng.normal(size=(s.shape[0] - 4, 3)) * 0.6
        dirs.append(np.asfortranarray(dd))
    out["self_dirs"] = np.array(dirs)
    out["self_steps"] = o.self_step(splines, dirs)
    out["couple_step"] = o.couple_self_step(splines, dirs)
    np.savez_compressed(os.path.join(HERE, "multi.npz"), **out)
    print("multi.npz: self planes", len(sd), "accepted hull pairs", int(np.sum(oks)), "self steps", out["self_steps"], out["couple_step"])


def coupled():
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    sc["way_points"] = sc["way_points"][:2] + sc["way_points"][4:6]
    U, P = 4, 4
    o = oa.RefOracle(); o.setup(oa.Params(P, uav_num=U, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in sc["way_points"]]
    out = {}
    for i in range(1, 7):
        sts = o.optimization_multi(sts, coupled=True)
        for u, s in enumerate(sts):
            for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
                out["it%d_u%d_%s" % (i, u, k)] = s[k]
        out["it%d_piece_time" % i] = sts[0]["piece_time"]; out["it%d_gnorm" % i] = sts[0]["gnorm"]
    np.savez_compressed(os.path.join(HERE, "coupled.npz"), **out)
    print("coupled.npz: piece_time", [float(out["it%d_piece_time" % i]) for i in range(1, 7)])


def optplane():
    out = {}
    # ---- single UAV, persistent obstacle planes
    sc = scenes.bridge(n_pts=4000, seed=21, n_pieces=4)
    P = 4
    o = oa.RefOracle(); o.setup(oa.Params(P, ks=sc["ks"], optimal_plane=1)); o.init_pointcloud(sc["V"]); o.reset_persistent_planes()
    st = scenes.initial_states(sc)[0]
    for i in range(1, 7):
        st = o.optimization(st)
        for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
            out["s_it%d_%s" % (i, k)] = st[k]
        out["s_it%d_piece_time" % i] = st["piece_time"]; out["s_it%d_gnorm" % i] = st["gnorm"]
        if i in (1, 3, 6):
            tr, ids, c, d = o.live_planes()
            out["s_it%d_live_tr" % i] = tr; out["s_it%d_live_id" % i] = ids; out["s_it%d_live_c" % i] = c; out["s_it%d_live_d" % i] = d
    # per-pair optimal_cd on real candidates of the initial trajectory, started from the GJK plane
    o.setup(oa.Params(P, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    sp = scenes.initial_states(sc)[0]["spline"]
    off, ids = o.dcd_collision(sp, 0.2)
    Pm = np.array([o.segment_points(sp, tr) for tr in range(P * 8)])
    rng = np.random.default_rng(11)
    pairs = [(tr, pid) for tr in range(P * 8) for pid in ids[off[tr]:off[tr + 1]]]
    pairs = [pairs[i] for i in rng.choice(len(pairs), size=min(3000, len(pairs)), replace=False)]
    rec = []
    for tr, pid in pairs:
        q = sc["V"][pid]
        ok, c, d = o.opengjk(Pm[tr], q.reshape(1, 3), 0.2)
        if ok:
            c1, d1 = o.optimal_cd(Pm[tr], q, c, d)
            rec.append((Pm[tr], q, c, d, c1, d1))
        if len(rec) >= 400:
            break
    for j, nm in enumerate(("cd_P", "cd_q", "cd_c0", "cd_d0", "cd_c1", "cd_d1")):
        out[nm] = np.array([r[j] for r in rec])
    # ---- 4 UAVs crossing, persistent inter-robot planes
    sc = scenes.cross(n_pts=3000, seed=23, n_pieces=4)
    sc["way_points"] = sc["way_points"][:2] + sc["way_points"][4:6]
    U = 4
    o.setup(oa.Params(P, uav_num=U, ks=sc["ks"], optimal_plane=1)); o.init_pointcloud(sc["V"]); o.reset_persistent_planes()
    sts = [scenes.init_state(scenes.init_spline_multi(wp)) for wp in sc["way_points"]]
    # self_optimal_cd on the hull pairs of the initial trajectories
    splines = [s["spline"] for s in sts]
    rec = []
    for tr in range(P * 8):
        Pl = [o.segment_points(s, tr) for s in splines]
        for a in range(U):
            for b in range(a + 1, U):
                ok, c, d = o.selfgjk(Pl[a], Pl[b], 0.3)
                if ok:
                    c1, d1 = o.self_optimal_cd(Pl[a], Pl[b], c, d)
                    rec.append((Pl[a], Pl[b], c, d, c1, d1))
    # ... and on random near-touching hull pairs.  self_optimal_cd shifts an indefinite 3x3 Hessian to a smallest eigenvalue
    # of 1e-8 and may take thousands of tiny steps, so part of the pairs are ill-conditioned (the result moves by 1e-3 when
    # an input moves by one ulp): `scd_stable` marks the pairs whose reference result is insensitive to such a perturbation;
    # only those are compared tightly, the others through the stopping criterion and the barrier energy.
    rng = np.random.default_rng(5)
    while len(rec) < 300:
        base = rng.uniform(-1, 1, 3)
        dirn = rng.normal(size=3); dirn /= np.linalg.norm(dirn)
        P0 = np.asfortranarray(base + np.outer(np.linspace(0, 0.6, 6), dirn) + rng.normal(scale=0.02, size=(6, 3)))
        dir2 = rng.normal(size=3); dir2 /= np.linalg.norm(dir2)
        sep = rng.normal(size=3); sep /= np.linalg.norm(sep)
        P1 = np.asfortranarray(base + sep * rng.uniform(0.12, 0.3) + np.outer(np.linspace(-0.3, 0.3, 6), dir2) + rng.normal(scale=0.02, size=(6, 3)))
        ok, c, d = o.selfgjk(P0, P1, 0.3)
        # feasible pairs only (hull distance > offset, what the inter-robot CCD step maintains): from an infeasible start
        # the reference's barrier is +inf / log of a negative number and the plane degenerates to NaN
        if not ok or not np.isfinite(c).all() or np.linalg.norm(o.gjk(P0, P1)) < 0.105:
            continue
        c1, d1 = o.self_optimal_cd(P0, P1, c, d)
        rec.append((P0, P1, c, d, c1, d1))
    stable = []
    for P0, P1, c, d, c1, d1 in rec:
        c2, d2 = o.self_optimal_cd(P0 * (1 + 2e-16), P1, c, d)
        stable.append(max(np.abs(c2 - c1).max(), abs(d2 - d1)) < 1e-9)
    out["scd_stable"] = np.array(stable)
    for j, nm in enumerate(("scd_P0", "scd_P1", "scd_c0", "scd_d0", "scd_c1", "scd_d1")):
        out[nm] = np.array([r[j] for r in rec])
    for i in range(1, 5):
        sts = o.optimization_multi(sts, coupled=False)
        for u, s in enumerate(sts):
            out["m_it%d_u%d_spline" % (i, u)] = s["spline"]
            out["m_it%d_u%d_piece_time" % (i, u)] = s["piece_time"]
        out["m_it%d_gnorm" % i] = sts[0]["gnorm"]
    np.savez_compressed(os.path.join(HERE, "optplane.npz"), **out)
    print("optplane.npz: live planes", len(out["s_it6_live_d"]), "cd pairs", len(out["cd_d1"]), "self pairs", len(out["scd_d1"]),
          "of which stable", int(np.sum(out["scd_stable"])))


if __name__ == "__main__":
    which = sys.argv[1:] or ["single", "multi", "coupled", "optplane"]
    for w in which:
        {"single": single, "multi": multi, "coupled": coupled, "optplane": optplane}[w]()
