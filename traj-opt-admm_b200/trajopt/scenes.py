"""Synthetic inputs of the named shapes; the Google-Drive dataset of the reference
(README.md:28) is unavailable offline.  All clouds are FP64, seeded, and feasible for the straight-line
initial trajectory (every obstacle point is farther than `offset` from the initial control hull and inside or
near the barrier band offset < dist < offset+margin so that planes are active from iteration 0).

Way-points -> control points follows the reference's two executables:
  single UAV : Main/admmPathPlanning3D.cpp:249-276   (init_spline_single)
  multi  UAV : Main/multiPathPlanning3D.cpp:342-375  (init_spline_multi)
"""
import numpy as np

ORDER = 5


def convert_list(piece_num):
    """C2-continuity conversion matrices with time_weight == 1 (CCDUtils.h:142-169): p = q = 0.5."""
    out = np.zeros((piece_num, 6, 6))
    for i in range(piece_num):
        out[i] = np.eye(6)
    p = q = 0.5
    I0 = np.array([[q * q, 2 * p * q, p * p], [0, q, p]])
    I1 = np.array([[q, p, 0], [q * q, 2 * p * q, p * p]])
    for i in range(piece_num - 1):
        out[i][4:6, 3:6] = I1
        out[i + 1][0:2, 0:3] = I0
    return out


def init_spline_single(way_points):
    wp = np.asarray(way_points, dtype=np.float64)
    P = wp.shape[0] - 1
    T = 6 + 3 * (P - 1)
    s = np.zeros((T, 3), order="F")
    s[0] = wp[0]
    for i in range(P):
        head = 0.9 * wp[i] + 0.1 * wp[i + 1]
        tail = 0.9 * wp[i + 1] + 0.1 * wp[i]
        s[3 * i + 1] = wp[i]
        s[3 * i + 2] = head
        s[3 * i + 3] = tail
        s[3 * (i + 1) + 1] = wp[i + 1]
    s[T - 1] = wp[P]
    s[1] = s[0]
    s[T - 2] = s[T - 1]
    return s


def init_spline_multi(way_points):
    wp = np.asarray(way_points, dtype=np.float64)
    P = wp.shape[0] - 1
    T = 6 + 3 * (P - 1)
    s = np.zeros((T, 3), order="F")
    s[0] = wp[0]
    for k in range(P):
        for j in range(0, 4):
            s[j + 3 * k + 1] = (3 - j) / 3.0 * wp[k] + j / 3.0 * wp[k + 1]
    s[T - 1] = wp[P]
    s[1] = s[0]
    s[T - 2] = s[T - 1]
    return s


def init_state(spline, piece_time=20.0):
    """slack / dual initialisation of init_variable (admmPathPlanning3D.cpp:278-292)."""
    T = spline.shape[0]
    P = (T - 6) // 3 + 1
    cv = convert_list(P)
    p_slack = np.zeros((6 * P, 3), order="F")
    for sp in range(P):
        # same accumulation order as Eigen's coefficient-based 6x6 * 6x3 product (sequential in k)
        blk = spline[3 * sp:3 * sp + 6]
        acc = np.zeros((6, 3))
        for k in range(6):
            acc = acc + cv[sp][:, k:k + 1] * blk[k:k + 1, :]
        p_slack[6 * sp:6 * sp + 6] = acc
    return dict(spline=np.array(spline, order="F"), piece_time=float(piece_time), p_slack=p_slack,
                t_slack=np.full(P, float(piece_time)), p_lambda=np.zeros((6 * P, 3), order="F"), t_lambda=np.zeros(P))


def straight_waypoints(a, b, n_pieces):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.array([a + (b - a) * (i / n_pieces) for i in range(n_pieces + 1)])


# ---------------------------------------------------------------------------------------------------------
def bridge(n_pts=100_000, seed=1, n_pieces=8):
    """C1: deck slab + two portal frames + side rails around the straight path (-7,0,0)->(7,0,0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_deck = int(0.4 * n_pts); n_frame = int(0.4 * n_pts); n_rail = n_pts - n_deck - n_frame
    deck = np.empty((n_deck, 3))
    deck[:, 0] = rng.uniform(-6, 6, n_deck); deck[:, 1] = rng.uniform(-1, 1, n_deck)
    deck[:, 2] = np.where(rng.random(n_deck) < 0.5, -0.15, -0.35)
    frame = np.empty((n_frame, 3))
    side = np.where(rng.random(n_frame) < 0.5, -2.0, 2.0)
    frame[:, 0] = side + rng.uniform(-0.05, 0.05, n_frame)
    part = rng.random(n_frame)
    col = part < 0.7
    frame[:, 1] = np.where(col, np.where(rng.random(n_frame) < 0.5, -0.18, 0.18), rng.uniform(-0.18, 0.18, n_frame))
    frame[:, 2] = np.where(col, rng.uniform(-0.15, 1.2, n_frame), 0.17)
    rail = np.empty((n_rail, 3))
    rail[:, 0] = rng.uniform(-6, 6, n_rail)
    rail[:, 1] = np.where(rng.random(n_rail) < 0.5, -0.45, 0.45)
    rail[:, 2] = rng.uniform(-0.15, 0.45, n_rail)
    V = np.concatenate([deck, frame, rail])
    V = V[rng.permutation(V.shape[0])]
    wp = straight_waypoints((-7, 0, 0), (7, 0, 0), n_pieces)
    return dict(name="bridge", V=np.asfortranarray(V), way_points=[wp], uav_num=1, ks=1e-8)


def _poisson_disk_centres(rng, n, xlim, ylim, min_dist, y_excl):
    pts = np.zeros((0, 2))
    cell = min_dist
    grid = {}
    out = []
    tries = 0
    while len(out) < n and tries < 400 * n:
        tries += 1
        x = rng.uniform(*xlim); y = rng.uniform(*ylim)
        if abs(y) < y_excl:
            continue
        gx, gy = int(np.floor(x / cell)), int(np.floor(y / cell))
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for (px, py) in grid.get((gx + dx, gy + dy), ()):
                    if (px - x) ** 2 + (py - y) ** 2 < min_dist ** 2:
                        ok = False
        if ok:
            grid.setdefault((gx, gy), []).append((x, y))
            out.append((x, y))
    return np.array(out)


def forest(n_pts=1_000_000, seed=2, n_pieces=64, n_trunks=640, half_len=32.0, half_width=1.0, ground_frac=0.25):
    """C2: a trail through a dense forest: vertical trunks (cylinders r=0.15, z in [-1,1]) at Poisson-disk
    positions in the corridor |y|<half_width around the straight path along x (centres kept out of |y|<0.28 so
    the nearest trunk surface is >= 0.13 from the path: feasible and inside the 0.2 barrier band), plus ground
    cover at z=-0.15 (also inside the band)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ctr = _poisson_disk_centres(rng, n_trunks, (-half_len, half_len), (-half_width, half_width), 0.36, 0.28)
    n_ground = int(ground_frac * n_pts)
    n_tr_pts = n_pts - n_ground
    per = n_tr_pts // len(ctr)
    rem = n_tr_pts - per * len(ctr)
    counts = np.full(len(ctr), per); counts[:rem] += 1
    idx = np.repeat(np.arange(len(ctr)), counts)
    ang = rng.uniform(0, 2 * np.pi, n_tr_pts)
    V = np.empty((n_pts, 3))
    V[:n_tr_pts, 0] = ctr[idx, 0] + 0.15 * np.cos(ang)
    V[:n_tr_pts, 1] = ctr[idx, 1] + 0.15 * np.sin(ang)
    V[:n_tr_pts, 2] = rng.uniform(-1, 1, n_tr_pts)
    V[n_tr_pts:, 0] = rng.uniform(-half_len, half_len, n_ground)
    V[n_tr_pts:, 1] = rng.uniform(-half_width, half_width, n_ground)
    V[n_tr_pts:, 2] = -0.15
    V = V[rng.permutation(n_pts)]
    wp = straight_waypoints((-half_len, 0, 0), (half_len, 0, 0), n_pieces)
    return dict(name="forest", V=np.asfortranarray(V), way_points=[wp], uav_num=1, ks=1e-8)


def tube(n_pts, seed, radius, n_pieces=8, half_len=7.0):
    """C5 member: cylinder of the given radius around the straight path (survey 8(d) 'tube clouds')."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ang = rng.uniform(0, 2 * np.pi, n_pts)
    V = np.empty((n_pts, 3))
    V[:, 0] = rng.uniform(-half_len + 1, half_len - 1, n_pts)
    V[:, 1] = radius * np.cos(ang)
    V[:, 2] = radius * np.sin(ang)
    wp = straight_waypoints((-half_len, 0, 0), (half_len, 0, 0), n_pieces)
    return dict(name="tube", V=np.asfortranarray(V), way_points=[wp], uav_num=1, ks=1e-8)


def cross(n_pts=50_000, seed=3, n_pieces=8):
    """C3: 8 UAVs, 4 lanes flying +x at z=0 and 4 lanes flying +y at z=0.25; floor + ceiling cloud.
    Coordinates are post-x5 (the multi executable scales its files by 5, multiPathPlanning3D.cpp:107,536)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    V = np.empty((n_pts, 3))
    V[:, 0] = rng.uniform(-6, 6, n_pts); V[:, 1] = rng.uniform(-6, 6, n_pts)
    V[:, 2] = np.where(rng.random(n_pts) < 0.5, -0.17, 0.42)
    wps = []
    for y in (-1.5, -0.5, 0.5, 1.5):
        wps.append(straight_waypoints((-5, y, 0), (5, y, 0), n_pieces))
    for x in (-1.5, -0.5, 0.5, 1.5):
        wps.append(straight_waypoints((x, -5, 0.25), (x, 5, 0.25), n_pieces))
    return dict(name="cross", V=np.asfortranarray(V), way_points=wps, uav_num=8, ks=1e-3)


def circle(n_uav=64, n_pts=20_000, seed=4, n_pieces=8, radius=10.0, eps=0.35, stagger=0.2):
    """C4: antipodal swap on a circle, goal = start rotated by pi-eps, altitude stagger `stagger`*(i mod 4) (a pure
    antipodal straight-line init is infeasible: all hulls meet at the centre).  With eps=0.35 all robots cross a ring
    of radius R*sin(eps/2)=1.74 at mid-flight, neighbours 0.17 apart in xy and 0.2 apart in z: inside the inter-robot
    activation distance offset+2*margin=0.3 and outside offset=0.1.  Obstacle ring at radius 12."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ang = rng.uniform(0, 2 * np.pi, n_pts)
    V = np.empty((n_pts, 3))
    V[:, 0] = 12.0 * np.cos(ang); V[:, 1] = 12.0 * np.sin(ang); V[:, 2] = rng.uniform(-0.5, 1.5, n_pts)
    wps = []
    for i in range(n_uav):
        a0 = 2 * np.pi * i / n_uav
        a1 = a0 + np.pi - eps
        z = stagger * (i % 4)
        wps.append(straight_waypoints((radius * np.cos(a0), radius * np.sin(a0), z),
                                      (radius * np.cos(a1), radius * np.sin(a1), z), n_pieces))
    return dict(name="circle", V=np.asfortranarray(V), way_points=wps, uav_num=n_uav, ks=1e-3)


def batch_member_meta(k):
    """(cloud size, tube radius) of problem k without generating the cloud"""
    rng = np.random.Generator(np.random.PCG64(1000 + k))
    n = int(round(10 ** rng.uniform(4, 6)))
    r = rng.uniform(0.13, 1.0)
    return n, r


def batch_cost(n, r, n_pieces=8, half_len=7.0, band=0.2):
    """expected cost of one ADMM iteration of a batch member, in arbitrary units, from its cloud size and tube radius alone
    (straight initial path along x): the obstacle candidates are the points within `band` of a sub-segment's box in y and z
    (a tube of radius r: all of them for r <= band, an arc of them up to r = band*sqrt(2), none beyond), each seen by the
    sub-segments whose x-range (+- band) covers it; the planes are the candidates within `band` of the segment itself.
    Weights: narrowphase per candidate, barrier energy / gradient / packing per plane, a fixed part per problem (measured
    on the 1024-problem batch: 0.15 ns per candidate, 0.22 ns per plane, 2.4 us per problem)"""
    seg = 2.0 * half_len / (8 * n_pieces)
    if r <= band:
        frac = 1.0
    elif r < band * np.sqrt(2.0):
        frac = (np.arcsin(band / r) - np.arccos(band / r)) / (np.pi / 2)
    else:
        frac = 0.0
    cand = n * frac * (seg + 2 * band) / seg
    planes = n * (seg + 2 * np.sqrt(band * band - r * r)) / seg if r < band else 0.0
    return 0.15 * cand + 0.22 * planes + 2400.0 + 0.002 * n


def batch_partition(total, world, rank):
    """problems of `rank`: longest-processing-time-first on the expected cost of a problem (batch_cost): the problems are taken
    in order of decreasing cost and each goes to the rank with the least cost so far (ties: lowest rank).  Deterministic,
    the same on every rank; returns this rank's problems in that order (heaviest first)."""
    meta = [batch_member_meta(k) for k in range(total)]
    cost = [batch_cost(n, r) for n, r in meta]
    order = sorted(range(total), key=lambda k: (-cost[k], k))
    load = [0.0] * world
    mine = []
    for k in order:
        owner = min(range(world), key=lambda w: (load[w], w))
        load[owner] += cost[k]
        if owner == rank:
            mine.append(k)
    return mine


def batch_member(k, n_pieces=8):
    """C5: problem k of the 1024-problem sweep: cloud size log-uniform in [1e4,1e6], tube radius in [0.13,1.0]."""
    n, r = batch_member_meta(k)
    sc = tube(n, 1000 + k, r, n_pieces)
    sc["name"] = "batch%d" % k
    return sc


def initial_states(scene, piece_time=20.0):
    single = scene["uav_num"] == 1
    f = init_spline_single if single else init_spline_multi
    return [init_state(f(wp), piece_time) for wp in scene["way_points"]]


def write_reference_files(scene, root, name="scene.obj", config=None):
    """Lay the scene out the way the reference executables read it (cwd = root):
      Config_File/3D.json                      Main/admmPathPlanning3D.cpp:368-397 (note the underscore)
      model/single/<name> | model/multiple/<name>   OBJ with `v x y z` lines (CCDUtils.h:317-391)
      init/<name>_init_file.txt                one way-point per line (single, :82-101); all robots side by side, 3 columns
                                               each (multi, multiPathPlanning3D.cpp:78-121)
      result/                                  result file directory
    The multi executable multiplies cloud and way-points by 5 (multiPathPlanning3D.cpp:107,536): files store /5."""
    import json
    import os
    multi = scene["uav_num"] > 1
    scale = 0.2 if multi else 1.0
    for d in ("Config_File", "model/single", "model/multiple", "init", "result"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    cfg = {"lambda": 10, "epsilon": 0.1, "margin": 0.1, "offset": 0.1, "res": 8, "vel_limit": 2, "acc_limit": 2, "mu": 0.1,
           "stop": 1e-2, "optimal_plane": 0, "decouple": 1, "init": 1, "init_ob": 1, "gui": 0, "exit": 0, "auto": 0}
    cfg.update(config or {})
    with open(os.path.join(root, "Config_File", "3D.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    V = np.asarray(scene["V"]) * scale
    with open(os.path.join(root, "model", "multiple" if multi else "single", name), "w") as f:
        for p in V:
            f.write("v %.17g %.17g %.17g\n" % (p[0], p[1], p[2]))
    wps = [np.asarray(w) * scale for w in scene["way_points"]]
    with open(os.path.join(root, "init", name + "_init_file.txt"), "w") as f:
        for i in range(len(wps[0])):
            f.write(" ".join("%.17g %.17g %.17g" % tuple(w[i]) for w in wps) + "\n")
    return cfg
