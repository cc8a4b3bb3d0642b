"""the reference's file formats (OBJ vertex reader with its quirks, init files, result file, trajectory
sampling / length) in trajopt/io.py against the compiled reference (Mesh::readOBJ through oracle/_ref) and against the output
of the reference executables.  CPU only."""
import os
import re
import subprocess

import numpy as np
import pytest

from trajopt import io as tio, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


OBJ_CASES = {
    "plain": "v 1 2 3\nv 4 5 6\nv -1.5e-3 0.25 7\n",
    "colours_and_comments_before": "# header\no mesh\nv 1 2 3 0.5 0.5 0.5\nv 4 5 6 1 1 1\n\nv 7 8 9\n",
    # more than 10 vertices, then a face: everything behind the first non-vertex line is never read
    "stops_after_vertex_block": "".join("v %d %d %d\n" % (i, 2 * i, 3 * i) for i in range(12)) + "f 1 2 3\nv 100 100 100\n",
    # only 5 vertices: a non-vertex line does NOT end the file yet
    "few_vertices_keep_reading": "".join("v %d 0 0\n" % i for i in range(5)) + "vn 0 0 1\nf 1 2 3\nv 9 9 9\n",
    "normals_are_not_vertices": "".join("v %d 1 1\n" % i for i in range(11)) + "vn 0 0 1\nv 5 5 5\n",
    "empty_lines_do_not_stop": "".join("v %d 1 1\n" % i for i in range(11)) + "\n\nv 5 5 5\n# end\nv 6 6 6\n",
    "tabs_and_spaces": "v\t1.0   2.0\t3.0\nv 1e2 -2E-2 +3\n",
}


@pytest.mark.parametrize("name", sorted(OBJ_CASES))
def test_obj_reader_matches_reference(oracle_ref, tmp_path, name):
    path = tmp_path / (name + ".obj")
    path.write_text(OBJ_CASES[name])
    ref = oracle_ref.read_obj(path)
    got = tio.read_obj_vertices(str(path))
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(got, ref)


def test_obj_roundtrip_of_a_scene_file(oracle_ref, tmp_path):
    sc = scenes.bridge(n_pts=3000, seed=2)
    scenes.write_reference_files(sc, str(tmp_path), "b.obj")
    p = os.path.join(str(tmp_path), "model", "single", "b.obj")
    ref = oracle_ref.read_obj(p)
    got = tio.read_obj_vertices(p)
    assert np.array_equal(ref, got) and np.array_equal(got, sc["V"])      # %.17g round-trips FP64


def test_init_files(tmp_path):
    sc = scenes.bridge(n_pts=100, seed=2)
    scenes.write_reference_files(sc, str(tmp_path), "s.obj")
    wp = tio.read_init_file(os.path.join(str(tmp_path), "init", "s.obj_init_file.txt"))
    assert np.array_equal(wp, sc["way_points"][0])
    scm = scenes.cross(n_pts=100, seed=3)
    scenes.write_reference_files(scm, str(tmp_path), "m.obj")
    wps = tio.read_init_file(os.path.join(str(tmp_path), "init", "m.obj_init_file.txt"), multi=True)
    assert len(wps) == scm["uav_num"]
    for a, b in zip(wps, scm["way_points"]):
        assert np.allclose(a, b, rtol=1e-15, atol=1e-15)       # stored /5, multiplied by 5 on reading
    # operator>> semantics on short / empty lines: first missing coordinate becomes 0, the rest keep the previous values
    p = tmp_path / "q_init_file.txt"
    p.write_text("1 2 3\n4 5\n\n7 8 9\n")
    assert np.array_equal(tio.read_init_file(str(p)), np.array([[1, 2, 3], [4, 5, 0], [0, 5, 0], [7, 8, 9.0]]))


def test_result_file_and_trajectory_length_match_the_reference_executable(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "admmPathPlanning3D_ref")
    if not os.path.exists(exe):
        pytest.skip("reference executable not built (needs /root/reference)")
    sc = scenes.bridge(n_pts=1500, seed=4)
    root = str(tmp_path)
    scenes.write_reference_files(sc, root, "b.obj", {"stop": 1e9, "exit": 1})     # stops at the first test of gnorm: iter == 2
    out = subprocess.run([exe, "b.obj"], cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    res = tio.read_result_file(os.path.join(root, "result", "b.obj_result_file_admm.txt"))
    assert res["iter"] == 2 and res["n_points"] == 1500 and res["running_time_ms"] >= 0
    tio.write_result_file(os.path.join(root, "mine.txt"), res["iter"], res["running_time_ms"], res["n_points"])
    assert tio.read_result_file(os.path.join(root, "mine.txt")) == res
    # "ccd time" / "ccd len" of log_data for the same trajectory: two iterations of the oracle from the same initial state
    from oracle import oracle_api as oa
    o = oa.get()
    o.setup(oa.Params(8, ks=sc["ks"])); o.init_pointcloud(sc["V"])
    st = scenes.initial_states(sc)[0]
    for _ in range(2):
        st = o.optimization(st)
    t_ref = float(re.search(r"^ccd time:([-+0-9.eE]+)", out.stdout, flags=re.M).group(1))
    l_ref = float(re.search(r"^ccd len:([-+0-9.eE]+)", out.stdout, flags=re.M).group(1))
    t, length = tio.trajectory_length(st["spline"], st["piece_time"])
    assert abs(t - t_ref) <= 1e-5 * t_ref and abs(length - l_ref) <= 1e-5 * l_ref      # cout prints 6 significant digits
