"""CPU side of the long-horizon study (tests/test_gpu_drift.py holds the GPU assertion): how reproducible is the reference
against ITSELF?  Runs here, without a GPU, on the compiled reference."""
import os

import numpy as np
import pytest

from trajopt import scenes
from oracle import oracle_api as oa

from test_gpu_drift import RefO2, run_ref, perturbed_states, dist


def test_reference_is_bitwise_stable_across_optimisation_levels(oracle_ref):
    """the -O2 build of the unmodified reference reproduces the -O3 build bit for bit: compiler optimisation level is not a
    source of drift for this code (no -ffast-math), one-ulp input changes are"""
    try:
        o2 = RefO2()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/O2 not built")
    sc = scenes.bridge(n_pts=6000, seed=5)
    st0 = scenes.initial_states(sc)
    a = run_ref(oracle_ref, sc, st0, 1e-2, max_it=40)
    b = run_ref(o2, sc, st0, 1e-2, max_it=40)
    assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_one_ulp_of_one_input_moves_the_reference_far_beyond_1e_6(oracle_ref):
    """the premise of the envelope test: the reference amplifies a one-ulp change of one control-point coordinate to > 1e-6
    within ~20 iterations and to > 1e-5 at convergence (bridge scene; it still stops at the same iteration)"""
    sc = scenes.bridge(n_pts=6000, seed=5)
    st0 = scenes.initial_states(sc)
    ref = run_ref(oracle_ref, sc, st0, 1e-2)
    pert = run_ref(oracle_ref, sc, perturbed_states(st0, 1, seed=1)[0], 1e-2)
    d = dist(ref, pert)
    assert d[0] < 1e-14 and d[:25].max() > 1e-6 and d[-1] > 1e-6
    assert abs(len(ref) - len(pert)) <= 1
