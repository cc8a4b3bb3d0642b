"""The two shortcuts of round 2 that rest on an argument rather than on an identity, checked against the COMPILED REFERENCE on
the CPU (the GPU tests check that the CUDA path gives the same result with and without them, tests/test_gpu_policy.py):

  * line search: a rung of the 0.8 ladder is not evaluated when a velocity / acceleration bound is violated there for certain
    (api.cu: k_ls_bound_mask).  Here: the reference's own Energy_admm::bound_energy is +inf at every rung the rule marks, on
    the states, directions and CCD steps of reference iterations.
  * narrowphase: axes 15..49 of the 49-DOP gate are evaluated only for pairs whose GJK distance is within 1e-6 (relative) of
    the gap (narrow.cu: k_narrow).  Here: for the broadphase candidates of reference runs, CCD::KDOPDCD passes whenever the
    reference's GJK witness is shorter than gap * (1 - 1e-6) -- the implication the shortcut uses -- and a pair the gate rejects
    never has a witness inside the gap.
"""
import numpy as np

from trajopt import scenes
from oracle import oracle_api as oa


def bound_mask(tab, p, st, direction, tdir, s0, n_rungs=32):
    """numpy restatement of k_ls_bound_mask: per rung, is some bound term below -1e-9 (limit + 1) ?"""
    x, t0 = st["spline"], st["piece_time"]
    mask = np.zeros(n_rungs, dtype=bool)
    s = s0
    for k in range(n_rungs):
        t = t0 + s * tdir
        xs = x + s * direction
        for tr in range(p.n_tr):
            B = tab["basis"][tr].reshape((6, 6), order="F")
            w = tab["weight"][tr]
            P = B @ xs[3 * (tr // p.res):3 * (tr // p.res) + 6, :]
            dv = p.vel_limit - 5 * np.linalg.norm(P[1:] - P[:-1], axis=1) / (w * t)
            da = p.acc_limit - 20 * np.linalg.norm(P[2:] - 2 * P[1:-1] + P[:-2], axis=1) / (w * w * t * t)
            if (dv < -1e-9 * (p.vel_limit + 1)).any() or (da < -1e-9 * (p.acc_limit + 1)).any():
                mask[k] = True
                break
        s *= 0.8
    return mask


def test_rungs_marked_by_the_bound_mask_are_infeasible_in_the_reference(oracle_ref):
    o = oracle_ref
    skipped = 0
    for sc in (scenes.tube(3000, 7, 0.3), scenes.bridge(n_pts=4000, seed=5), scenes.tube(2500, 11, 0.18)):
        P = len(sc["way_points"][0]) - 1
        p = oa.Params(P, ks=sc["ks"])
        o.setup(p)
        o.init_pointcloud(sc["V"])
        tab = o.tables()
        st = scenes.initial_states(sc)[0]
        for it in range(7):
            planes = o.separate_plane(st["spline"])
            direction, tdir, wolfe, gn = o.descent_direction(st, planes)
            s0 = o.position_step(st["spline"], direction)
            if st["piece_time"] + s0 * tdir <= 0:
                s0 = -0.95 * st["piece_time"] / tdir            # Optimization3D_admm.h:521-524
            mask = bound_mask(tab, p, st, direction, tdir, s0)
            s = s0
            ref_inf = np.zeros(32, dtype=bool)
            for k in range(32):
                ref_inf[k] = not np.isfinite(o.bound_energy(st["spline"] + s * direction, st["piece_time"] + s * tdir))
                s *= 0.8
            assert not (mask & ~ref_inf).any(), (it, np.nonzero(mask & ~ref_inf)[0])       # marked => +inf in the reference
            lead = 32 if mask.all() else int(np.argmin(mask))
            lead_ref = 32 if ref_inf.all() else int(np.argmin(ref_inf))
            assert lead <= lead_ref and lead_ref - lead <= 1, (it, lead, lead_ref)          # and nearly all of them are marked
            skipped += lead
            st = o.optimization(st)
    assert skipped > 40           # the rule is not vacuous on these runs


def test_gjk_witness_inside_the_gap_implies_the_kdop_gate(oracle_ref):
    o = oracle_ref
    rng = np.random.default_rng(3)
    n_acc = n_band = n_rej_gate = 0
    for sc in (scenes.bridge(n_pts=6000, seed=5), scenes.tube(4000, 9, 0.22)):
        P = len(sc["way_points"][0]) - 1
        p = oa.Params(P, ks=sc["ks"])
        o.setup(p)
        o.init_pointcloud(sc["V"])
        gap = p.offset + p.margin
        st = scenes.initial_states(sc)[0]
        for it in range(3):
            off, idx = o.dcd_collision(st["spline"], gap)
            rows = np.repeat(np.arange(len(off) - 1), np.diff(off))
            pick = rng.choice(len(idx), size=min(len(idx), 2500), replace=False)
            for j in pick:
                Pm = o.segment_points(st["spline"], int(rows[j]))
                q = sc["V"][int(idx[j])]
                w = o.gjk(Pm, q)
                cn = float(np.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]))
                gate = o.kdop_dcd(Pm, q, gap)
                if cn <= gap * (1.0 - 1e-6):
                    n_acc += 1
                    assert gate, (it, int(rows[j]), int(idx[j]), cn)         # the implication the shortcut relies on
                elif cn <= gap:
                    n_band += 1                                                # here the CUDA path evaluates the whole gate
                if not gate:
                    n_rej_gate += 1
                    assert cn > gap, (it, int(rows[j]), int(idx[j]), cn)      # a rejected pair never had a plane coming
            st = o.optimization(st)
    assert n_acc > 2000 and n_rej_gate > 100
    assert n_band <= 2
