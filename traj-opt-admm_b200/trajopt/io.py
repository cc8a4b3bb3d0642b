"""File formats of the reference's two executables, with their quirks, for hosts that feed the C ABI
without going through Main/*.cpp (bench.py, batch drivers).  Host-side, one-time per problem: plain Python.

  read_obj_vertices   Mesh::readOBJ, HighOrderCCD/Utils/CCDUtils.h:317-391
  read_init_file      way_point_init, Main/admmPathPlanning3D.cpp:79-112 (single) / Main/multiPathPlanning3D.cpp:78-121 (multi)
  write_result_file / read_result_file   Main/admmPathPlanning3D.cpp:400,507-510 ; multiPathPlanning3D.cpp (…_result_file_multi.txt)
  sample_trajectory / trajectory_length  log_data + getPosFromBezier, Main/admmPathPlanning3D.cpp:20-77 ("ccd time" / "ccd len")
"""
import math
import re

import numpy as np

from . import scenes

LINE_MAX = 2048      # IGL_LINE_MAX: fgets() splits longer lines


def _leading_doubles(text):
    """std::istream_iterator<double> over `text`: whitespace-separated numbers until the first token that is not one"""
    out = []
    for tok in text.split():
        try:
            out.append(float(tok))
        except ValueError:
            # operator>> consumes the longest numeric prefix of the token ("1.5abc" -> 1.5, then fails on "abc")
            m = re.match(r"[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", tok)
            if m:
                out.append(float(m.group(0)))
            break
    return out


def read_obj_vertices(path):
    """V (n x 3, column-major) exactly as Mesh::readOBJ builds it:
      * a line is a vertex iff its first whitespace-delimited word is exactly "v" ("vn", "vt", "f", "#..." are other lines);
      * the coordinates are parsed from the SECOND CHARACTER of the line on (&line[1]), so the "v" must be in column 0;
      * other non-empty lines are ignored, but once MORE THAN 10 vertices have been read the first such line ENDS the file
        (faces, normals, comments after the vertex block -- and any vertex behind them -- are never read);
      * empty lines are skipped without ending anything; extra numbers on a vertex line (colours) are ignored."""
    verts = []
    with open(path, "r", errors="replace") as f:
        data = f.read()
    for raw in data.split("\n"):
        for k in range(0, max(len(raw), 1), LINE_MAX - 1):      # fgets chunks of at most 2047 characters
            line = raw[k:k + LINE_MAX - 1]
            words = line.split()
            if not words:
                continue
            if words[0] == "v":
                xyz = _leading_doubles(line[1:])
                if len(xyz) < 3:
                    raise ValueError("readOBJ: vertex line without 3 coordinates (the reference reads out of bounds here): %r" % line[:60])
                verts.append(xyz[:3])
            elif len(verts) > 10:
                return np.asfortranarray(np.array(verts, dtype=np.float64).reshape(-1, 3))
    return np.asfortranarray(np.array(verts, dtype=np.float64).reshape(-1, 3))


def read_init_file(path, multi=False):
    """way-points of the initial path.  single: one `x y z` per line -> (n, 3).  multi: every line holds 3 columns per robot,
    uav_num = (words of the first line) / 3, and every point is MULTIPLIED BY 5 (multiPathPlanning3D.cpp:107) -> list of (n, 3).
    Like operator>> on a failed extraction, a short or empty line yields 0 for the first missing coordinate and keeps the
    previous values for the rest."""
    lines = open(path, "r").read().split("\n")
    if lines and lines[-1] == "":
        lines = lines[:-1]                       # getline() does not produce a line after the final newline
    if not multi:
        p = [0.0, 0.0, 0.0]
        out = []
        for ln in lines:
            vals = _leading_doubles(ln)
            for k in range(3):
                if k < len(vals):
                    p[k] = vals[k]
                else:
                    if k == len(vals):
                        p[k] = 0.0
                    break
            out.append(list(p))
        return np.array(out, dtype=np.float64).reshape(-1, 3)
    uav = len(lines[0].split()) // 3 if lines else 0
    way = [[] for _ in range(uav)]
    p = [0.0, 0.0, 0.0]
    for ln in lines:
        vals = _leading_doubles(ln)
        pos, failed = 0, False
        for j in range(uav):
            for k in range(3):
                if not failed and pos < len(vals):
                    p[k] = vals[pos]; pos += 1
                elif not failed:
                    p[k] = 0.0; failed = True
            p = [5.0 * x for x in p]             # p0 *= 5 acts on the running value, also after a failed read
            way[j].append(list(p))
    return [np.array(w, dtype=np.float64).reshape(-1, 3) for w in way]


def write_result_file(path, iters, running_time_ms, n_points):
    with open(path, "w") as f:
        f.write("iter: %d\nrunning time: %s\npoint cloud size: %d\n" % (iters, _cout(running_time_ms), n_points))


def read_result_file(path):
    txt = open(path).read()
    out = {}
    for key, name in (("iter", "iter"), ("running time", "running_time_ms"), ("point cloud size", "n_points")):
        m = re.search(r"^%s: ([-+0-9.eE]+)" % key, txt, flags=re.M)
        if m:
            out[name] = float(m.group(1)) if name == "running_time_ms" else int(float(m.group(1)))
    return out


def _cout(x):
    """std::ostream default formatting of a double (%g with 6 significant digits)"""
    return "%g" % x


def sample_trajectory(spline, piece_time, dt=0.05):
    """log_data: positions at t = 0, dt/piece_time, ... < piece_num (curve parameter, piece i = floor(t)), evaluated in the
    Bernstein basis of the C2-converted control points (convert_list[i] * block), as getPosFromBezier does"""
    spline = np.asarray(spline, dtype=np.float64)
    T = spline.shape[0]
    P = (T - 6) // 3 + 1
    cv = scenes.convert_list(P)
    comb = [math.comb(5, j) for j in range(6)]
    pts = []
    t = 0.0
    while t < P:
        i = int(math.floor(t))
        s = t - i
        bz = cv[i] @ spline[3 * i:3 * i + 6]
        pos = np.zeros(3)
        for ax in range(3):
            acc = 0.0
            for j in range(6):
                acc += comb[j] * bz[j, ax] * s ** j * (1 - s) ** (5 - j)
            pos[ax] = acc
        pts.append(pos)
        t += dt / piece_time
    return np.array(pts)


def trajectory_length(spline, piece_time):
    """("ccd time", "ccd len") of log_data: total duration piece_num * piece_time (time_weight == 1) and the polyline length"""
    pts = sample_trajectory(spline, piece_time)
    P = (np.asarray(spline).shape[0] - 6) // 3 + 1
    return P * piece_time, float(np.sum(np.linalg.norm(np.diff(pts, axis=0), axis=1)))
