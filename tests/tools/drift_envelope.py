"""Diagnostic (CPU only): how far does the compiled reference move from ITSELF when one input changes by one ulp?
Prints, per iteration to convergence, max|dspline| between the reference and runs of the same reference with one
control-point coordinate moved by one ulp, and the stopping iterations.  The yardstick of tests/test_gpu_drift.py.

    python tests/tools/drift_envelope.py bridge 1e-2
    python tests/tools/drift_envelope.py cross 1.0
    python tests/tools/drift_envelope.py cross 0.5 coupled
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from trajopt import scenes  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402
from test_gpu_drift import run_ref, perturbed_states, dist  # noqa: E402


def main():
    which, stop = sys.argv[1], float(sys.argv[2])
    coupled = len(sys.argv) > 3 and sys.argv[3] == "coupled"
    sc = scenes.bridge(n_pts=6000, seed=5) if which == "bridge" else scenes.cross(n_pts=4000, seed=3)
    st0 = scenes.initial_states(sc)
    o = oa.RefOracle()
    ref = run_ref(o, sc, st0, stop, coupled)
    perts = [run_ref(o, sc, stp, stop, coupled) for stp in perturbed_states(st0, 4, seed=1)]
    n = min([len(ref)] + [len(p) for p in perts])
    print("# %s%s: reference stops at iteration %d, one-ulp-perturbed runs at %s" % (which, " coupled" if coupled else "", len(ref) - 1, [len(p) - 1 for p in perts]))
    ds = [dist(ref, p) for p in perts]
    for i in range(n):
        print("it %3d  " % i + "  ".join("%.3e" % d[i] for d in ds))
    print("# final-vs-final: " + "  ".join("%.3e" % float(np.max(np.abs(ref[-1] - p[-1]))) for p in perts))


if __name__ == "__main__":
    main()
