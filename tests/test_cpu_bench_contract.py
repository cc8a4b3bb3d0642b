"""bench.py: what can be checked without a GPU -- the reference arm (the compiled reference on the host cores) prints one JSON
line with the keys the driver reads, for a single-UAV scene and for the batch (one oracle process per core on a bounded
sample); the work model of the kernels is sane; the GPU arm fails loudly (non-zero exit, no JSON line) when there is no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e")


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def _line(out):
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-2000:]      # ONE line on stdout, whatever libraries print
    return json.loads(lines[0])


@pytest.mark.parametrize("args", [("--workload", "bridge", "--steps", "2", "--warmup", "1"),
                                  ("--problems", "16", "--steps", "1", "--warmup", "1")])
def test_reference_arm_prints_the_contract_line(oracle_ref, args):
    out = _run("--impl", "reference", *args)
    assert out.returncode == 0, out.stderr[-2000:]
    j = _line(out)
    for k in NEED:
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "ADMM iters/sec" and j["unit"] == "iter/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and abs(j["ms_per_step"] * j["value"] - 1e3) < 1e-6 * 1e3
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert set(j["config"]) == {"workload", "l2", "multi_gpu"} and j["vs_baseline"] is None and j["dtype"] == "f64"


def test_work_model_of_the_kernels():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(b)
    finally:
        sys.argv = argv
    ps = {"dcd_candidates": 72.5e6, "ccd_candidates": 3.7e5, "planes": 40.7e6, "energy_plane_evals": 90e6, "barrier_terms": 4e8,
          "np_gjk_iters": 1.1e8, "np_kdop_groups": 1.2e8, "np_kdop_exact": 0.0, "ccd_kdop_pass": 0.0, "ccd_gjk_iters": 0.0}
    m = b.kernel_models(ps, {"rows": 65536, "P": 8, "T": 28, "U": 1024})
    for name in ("k_bp_count", "k_bp_fill", "k_narrow", "k_pack", "k_row_energy", "k_row_grad", "k_piece", "k_solve_bcr", "k_bp_ccd"):
        flop, byts, bound = m[name]
        assert flop >= 0 and byts > 0 and bound in ("hbm", "fp64"), name
    # counted work, not a worst case per candidate: GJK rounds + the gate groups really evaluated
    assert abs(m["k_narrow"][0] - (b.FLOP_GJK61_ROUND * 1.1e8 + b.FLOP_PLANE_FINISH * 40.7e6 + b.FLOP_KDOP_GROUP * 1.2e8)) < 1e-6 * m["k_narrow"][0]
    assert m["k_bp_fill"][1] < m["k_bp_count"][1]          # the fill pass scatters from records, it does not walk the tree


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    out = _run("--workload", "bridge", "--steps", "1", "--warmup", "1", "--no-cpu", timeout=300)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
