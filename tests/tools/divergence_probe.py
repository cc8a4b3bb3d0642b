"""Diagnostic (GPU box): per-iteration max |spline difference| between the compiled reference and the CUDA path on the
same scene, to tell round-off amplification (smooth growth) from a discrete decision flip (a jump)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)
from trajopt import api, scenes  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cross"
    coupled = len(sys.argv) > 2 and sys.argv[2] == "coupled"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 70
    sc = scenes.cross(n_pts=4000, seed=3) if which == "cross" else scenes.bridge(n_pts=6000, seed=5)
    U = sc["uav_num"]
    P = len(sc["way_points"][0]) - 1
    o = oa.get(); o.setup(oa.Params(P, uav_num=U, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    s = api.Solver(P, uav_num=U, ks=sc["ks"]); s.init_pointcloud(sc["V"])
    a = b = scenes.initial_states(sc)
    for it in range(iters):
        if U == 1:
            a = [o.optimization(a[0])]; b = [s.optimization(b[0])]
        else:
            a = o.optimization_multi(a, coupled=coupled); b = s.optimization(b, coupled=coupled)
        d = max(float(np.max(np.abs(x["spline"] - y["spline"]))) for x, y in zip(a, b))
        dt = max(abs(x["piece_time"] - y["piece_time"]) for x, y in zip(a, b))
        print("it %3d  max|dspline| %.3e  max|dt| %.3e  gnorm ref %.6g dev %.6g" % (it, d, dt, a[0]["gnorm"], b[0]["gnorm"]), flush=True)


if __name__ == "__main__":
    main()
