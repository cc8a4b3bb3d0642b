#!/usr/bin/env python
"""bench.py -- ADMM iterations/s (and segment-point pair evaluations/s) of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload batch|forest|bridge|cross8|circle64|circle64c] [--problems M]

Default workload = BASELINE.json configs[4], the configuration the metric is quoted on at 1/2/4/8 GPUs: the batched sweep
of 1024 independent single-UAV problems (tube clouds of 1e4..1e6 points, 214.6 M points in total, 8 Bezier pieces = 64
sub-segments each, FP64, Config File/3D.json parameters, straight-line initial trajectories).  One "step" = one ADMM
iteration (Optimization3D_admm::optimization, Optimization3D_admm.h:29-67) of EVERY problem; the fixed set of problems is
dealt over the ranks (strong scaling, no communication: SURVEY.md 8(e)).  --workload forest is configs[1] (one UAV, 1 M
points, 64 pieces: latency regime, replicas only on N > 1), circle64 / circle64c are configs[3] (64 UAVs sharded over the
ranks, decoupled / coupled, native NCCL exchange inside the iteration's CUDA graph), cross8 is configs[2].

  value  : problem-iterations/s (iterations/s for the single-problem workloads) with the state resident in HBM
           (tob_admm_iterate), CUDA events on the library's stream, L2 flushed between iterations (a 256 MiB buffer is
           rewritten outside the timed brackets), iterations W .. W+K-1 from the initial state.
  e2e    : the same iterations through the reference-shaped entry point with HOST buffers (tob_optimization = H2D of every
           problem's state + one iteration + D2H per call), wall clock around the call.
  --impl reference : the reference's own CPU implementation (oracle/_ref = the unmodified sources compiled here, else the
           C port) on the box's host cores, rank 0 only.  Batch: the reference is single-threaded and non-reentrant
           (globals), so independent oracle PROCESSES run a stratified sample of the problems concurrently, one per core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ADMM iters/sec"
UNIT = "iter/s"

# ---- algorithmic work per unit.  The pair kernels are costed from COUNTED work (device counters: 49-DOP groups evaluated,
# GJK rounds run, barrier terms inside the band), not from a worst-case constant per candidate; per-unit flop figures:
FLOP_KDOP_GROUP = 7 * (5 + 2)           # one group of 7 axes against precomputed extents: level (3 mul 2 add) + 2 subtractions.
                                        # FP64 in the default narrowphase variant; with the single-precision filter of the gate
                                        # (TRAJOPT_B200_NP_PMEM=0: gjk.cuh kdop_point_gate) reported as fp32 work instead
FLOP_KDOP_EXACT_AXIS = 7                # an axis the filter could not decide, re-tested in FP64
FLOP_KDOP_SHIFT = 3                     # per gate call: the point relative to the row's centre (3 FP64 subtractions)
FLOP_GJK61_ROUND = 45 + 120             # support over 6 points (30) + tests (15) + signed-volume sub-algorithm (S1D 30 / S2D 130 / S3D 330)
FLOP_GJK121_ROUND = 75 + 120            # same with 12 swept points
FLOP_PLANE_FINISH = 30                  # norm, normalise, d
FLOP_KDOP_SWEPT = 49 * (13 * 5 + 4)     # swept 49-DOP of a candidate that passes (general 12+1 point version)
FLOP_PER_PLANE_EVAL = 36.0              # 6 control points x (3 mul + 3 add)
FLOP_PER_ACTIVE_TERM_E = 20.0           # energy: 2 sub, 3 mul, 1 add, table logarithm 13 (csrc/fastlog.cuh; the library log was ~35)
FLOP_PER_ACTIVE_TERM_G = 75.0           # gradient: logarithm 13, reciprocal ~8, e1 / e2 ~30, 9 accumulates x 2, products of c
BYTES_PER_PLANE = 32.0
BYTES_PER_BUILD_POINT = 128.0           # SURVEY 8(d) U-build


# the gate of the narrowphase runs in FP64 unless the register variant with the single-precision filter is selected
GATE_FP32 = os.environ.get("TRAJOPT_B200_NP_PMEM") == "0" and os.environ.get("TRAJOPT_B200_NP_FILTER", "1") != "0"


def kernel_models(per_step, geo):
    """kernel -> (counted FP64 flop per step, SURVEY algorithmic bytes per step, bound).  Bytes follow SURVEY 8(d):
    U-bp = 48 B per row + 28 B per candidate (the level-1 pass over (row x node) boxes is served from L1/L2 and is NOT
    counted), U-np = 28 B read per candidate + 32 B per accepted plane."""
    rows, P, T, U = geo["rows"], geo["P"], geo["T"], geo["U"]
    cand, ccd, planes = per_step["dcd_candidates"], per_step["ccd_candidates"], per_step["planes"]
    evals, terms = per_step["energy_plane_evals"], per_step["barrier_terms"]
    e_evals = max(evals - planes, 0.0)                      # line-search passes (the gradient pass streams each plane once)
    g_share = planes / evals if evals else 0.0
    n_sys = 3 * (T - 4) + 1
    np_flop = (FLOP_GJK61_ROUND * per_step["np_gjk_iters"] + FLOP_PLANE_FINISH * planes
               + FLOP_KDOP_EXACT_AXIS * per_step.get("np_kdop_exact", 0.0) + FLOP_KDOP_SHIFT * 2 * cand)
    if not GATE_FP32:
        np_flop += FLOP_KDOP_GROUP * per_step.get("np_kdop_groups", 0.0) - FLOP_KDOP_SHIFT * 2 * cand
    ccd_flop = FLOP_KDOP_SWEPT * per_step["ccd_kdop_pass"] + FLOP_GJK121_ROUND * per_step["ccd_gjk_iters"]
    return {
        "k_rows": (0.0, rows * (18 + 6 + 2 * 49) * 8.0 * 2, "hbm"),
        "k_bp_count": (0.0, 48.0 * rows + 24.0 * cand, "hbm"),
        "k_bp_fill": (0.0, 4.0 * rows + 8.0 * cand, "hbm"),      # scatter from the count pass's records: (point, row) per candidate
                                                                  # written (+ 12 B per item record read, not counted: items are not)
        "k_bp_ccd": (ccd_flop, 48.0 * rows + 24.0 * ccd, "hbm"),
        "k_narrow": (np_flop, 28.0 * cand + 32.0 * planes, "fp64"),
        "k_pack": (0.0, 4.0 * cand + 72.0 * planes, "hbm"),
        "k_row_energy": (FLOP_PER_PLANE_EVAL * e_evals + FLOP_PER_ACTIVE_TERM_E * terms * (1 - g_share), BYTES_PER_PLANE * e_evals, "fp64"),
        "k_row_grad": (FLOP_PER_PLANE_EVAL * planes + FLOP_PER_ACTIVE_TERM_G * terms * g_share, BYTES_PER_PLANE * planes, "fp64"),
        "k_piece": (6.0 * (18 + 171) * 2 * rows + U * P * 19 ** 3 / 3.0, U * P * (361 + 19) * 8.0, "fp64"),
        "k_solve_bcr": (U * (n_sys * (17 ** 2 + 17) + 2 * n_sys * 17), U * P * (361 + 19) * 8.0, "fp64"),
        "k_slack": (U * P * (19 ** 3 / 3.0 + 4 * 19 * 19), U * P * (4 * 19) * 8.0, "fp64"),
        "k_robot_ls": (0.0, 9.0 * rows * 20, "hbm"),
    }


# ---- workloads ------------------------------------------------------------------------------------------------------
def batch_order(total):
    """the order batch_partition deals the problems in (expected pair work, then cloud size)"""
    from trajopt import scenes
    meta = [scenes.batch_member_meta(k) for k in range(total)]
    cost = [scenes.batch_cost(n, r) for n, r in meta]
    return sorted(range(total), key=lambda k: (-cost[k], k)), meta


def describe(args, world):
    """the `config` object: a pure function of the command line and the world size, identical in both arms"""
    from trajopt import scenes
    name = args.workload
    if name == "batch":
        total = args.problems or 1024
        pts = sum(scenes.batch_member_meta(k)[0] for k in range(total))
        wl = ("batch: %d independent single-UAV problems, tube clouds 1e4..1e6 pts (%.1f M pts in total), 8 Bezier pieces "
              "(64 sub-segments) each, 3D.json params" % (total, pts / 1e6))
        mg = "single" if world == 1 else "independent problems dealt over the ranks by expected cost (longest first to the least loaded rank), no communication"
    else:
        shape = {"forest": (1, args.points or 1_000_000, 64), "bridge": (1, args.points or 100_000, 8),
                 "cross8": (8, args.points or 50_000, 8), "circle64": (64, args.points or 20_000, 8),
                 "circle64c": (64, args.points or 20_000, 8)}[name]
        wl = "%s: %d UAV, %d pts, %d Bezier pieces (%d sub-segments each), 3D.json params" % ((name,) + shape + (shape[2] * 8,))
        if name == "circle64c":
            wl += ", coupled (decouple=0)"
        if shape[0] > 1 and world > 1:
            mg = "robots sharded over the ranks, cloud replicated, native NCCL all-gathers inside the iteration's CUDA graph"
        else:
            mg = "single" if world == 1 else "replicas only"
    return {"workload": wl, "l2": "flushed between timed iterations (256 MiB rewrite outside the event brackets)", "multi_gpu": mg}


def workload(args, rank, world):
    from trajopt import scenes
    name, n_pts = args.workload, args.points
    if name == "forest":
        return scenes.forest(n_pts=n_pts or 1_000_000)
    if name == "bridge":
        return scenes.bridge(n_pts=n_pts or 100_000)
    if name in ("circle64", "circle64c"):      # BASELINE.json configs[3]
        return scenes.circle(n_uav=64, n_pts=n_pts or 20_000)
    if name == "cross8":                       # configs[2]
        return scenes.cross(n_pts=n_pts or 50_000)
    if name == "batch":                        # configs[4]
        total = args.problems or 1024
        if args.emulate_rank:                  # profiling aid: the share rank r of w would get, on one GPU (never a bench value)
            r, w = (int(x) for x in args.emulate_rank.split("/"))
            mine = scenes.batch_partition(total, w, r)
        else:
            mine = scenes.batch_partition(total, world, rank)
        ms = [scenes.batch_member(k) for k in mine]
        return dict(name="batch", Vs=[m["V"] for m in ms], way_points=[m["way_points"][0] for m in ms], uav_num=len(ms), ks=1e-8,
                    n_points=sum(m["V"].shape[0] for m in ms), n_total=total)
    raise SystemExit("unknown workload " + name)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of the GPU this rank drives, sampled in-process through NVML every few ms DURING the timed
    region (an `nvidia-smi -lms` child needs ~1 s to start and initialises the driver inside the timed region).  Falls back
    to one `nvidia-smi` query per sample when pynvml is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, cuda_index, period=0.004):
        super().__init__(daemon=True)
        self.period, self.sm, self.reasons, self.smax, self.h, self.nv = period, [], set(), None, None, None
        self._stop_evt, self._on, self._hold = threading.Event(), threading.Event(), threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            for nm, bit in self.REASONS:
                if r & bit:
                    self.reasons.add(nm)
        else:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i",
                                  os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.sm.append(float(f[0])); self.smax = float(f[1])
            for (nm, _), v in zip(self.REASONS, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def run(self):
        while not self._stop_evt.is_set():
            if self._on.is_set() and not self._hold.is_set():
                try:
                    self._sample()
                except Exception:
                    pass
            time.sleep(self.period)

    # NVML queries go through the driver and can delay a CUDA launch issued at the same moment (measured: +12 % on a 0.35 ms
    # iteration when several ranks poll on one box): the main thread holds the sampler off while it is inside an event
    # bracket; samples are taken during the rest of the timed region (L2 flush + synchronise, GPU busy).
    def hold(self):
        self._hold.set()

    def release(self):
        self._hold.clear()

    def begin(self):
        self._on.set()

    def end(self):
        self._on.clear()

    def close(self):
        self._stop_evt.set()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---- CPU side: the reference on the host cores ----------------------------------------------------------------------
def _cpu_batch_worker(job):
    """one oracle process: its problems one after the other (the reference is non-reentrant: one problem per process)"""
    ks, warm, steps, budget = job
    sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)
    from oracle import oracle_api as oa
    from trajopt import scenes
    o = oa.get()
    out = []
    for k in ks:
        m = scenes.batch_member(k)
        o.setup(oa.Params(len(m["way_points"][0]) - 1, ks=m["ks"]))
        t0 = time.perf_counter()
        o.init_pointcloud(m["V"])
        build = time.perf_counter() - t0
        st = scenes.init_state(scenes.init_spline_single(m["way_points"][0]))
        for _ in range(warm):
            st = o.optimization(st)
        n, t = 0, 0.0
        while n < steps and (n == 0 or t < budget):
            t0 = time.perf_counter()
            st = o.optimization(st)
            t += time.perf_counter() - t0
            n += 1
        out.append((k, n, t, build, m["V"].shape[0]))
    return o.kind, out


def cpu_batch(total, warm, steps, budget_s, max_procs=128):
    """aggregate problem-iterations/s of `procs` concurrent oracle processes on a stratified sample of the batch (every
    (total/procs)-th problem of the work-sorted order, so the sample has the mix of the whole set).  With per-iteration
    times tau_k measured under that concurrency, one iteration of all problems on `procs` cores takes (total/S) sum tau_k /
    procs, i.e. the job runs at procs * S / sum tau_k problem-iterations/s."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, max_procs, total))
    order, _ = batch_order(total)
    per = 1 if procs >= 32 else -(-32 // procs)              # few cores: several problems per process, still >= 32 samples
    n_s = min(total, procs * per)
    sample = [order[int((i + 0.5) * total / n_s)] for i in range(n_s)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_batch_worker, [(sample[i::procs], warm, steps, budget_s / per) for i in range(procs)], chunksize=1)
    wall = time.perf_counter() - t0
    kind = res[0][0]
    rows = [r for _, lst in res for r in lst]
    tau = [t / n for _, n, t, _, _ in rows]
    value = procs * len(rows) / sum(tau)
    its = sorted(n for _, n, _, _, _ in rows)
    text = ("%d concurrent oracle processes (%d host cores visible), stratified sample of %d of the %d problems "
            "(clouds %d..%d pts); per problem %d untimed + %d..%d timed ADMM iterations from the initial state (time-bounded), "
            "tree builds excluded (%.0f s summed); wall %.0f s"
            % (procs, cores, len(rows), total, min(r[4] for r in rows), max(r[4] for r in rows), warm, its[0], its[-1],
               sum(r[3] for r in rows), wall))
    return {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": text}


def cpu_single(sc, P, budget_s, max_iters, warm=0):
    """the reference's CPU path on ONE scene: 1 core (the reference has no parallel region and is non-reentrant)"""
    from oracle import oracle_api as oa
    from trajopt import scenes
    o = oa.get()
    o.setup(oa.Params(P, uav_num=sc["uav_num"], ks=sc["ks"]))
    t0 = time.time()
    o.init_pointcloud(sc["V"])
    build = time.time() - t0
    sts = scenes.initial_states(sc)
    coupled = sc.get("coupled", False)
    step = (lambda x: o.optimization_multi(x, coupled=coupled)) if len(sts) > 1 else (lambda x: [o.optimization(x[0])])
    t_w = 0.0
    for _ in range(warm):
        t0 = time.perf_counter(); sts = step(sts); t_w += time.perf_counter() - t0
        if t_w > budget_s / 3:
            break
    n, t_used = 0, 0.0
    while n < max_iters and (n == 0 or t_used < budget_s):
        t0 = time.perf_counter()
        sts = step(sts)
        t_used += time.perf_counter() - t0
        n += 1
    return {"value": n / t_used, "unit": UNIT, "cores": 1, "kind": o.kind,
            "sample": "%d ADMM iterations of the same scene from the initial state after %d untimed ones, 1 host core "
                      "(tree build %.1f s excluded)" % (n, warm, build)}


# ---- our arm --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from trajopt import api, scenes

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    sc = workload(args, rank, world)
    P = len(sc["way_points"][0]) - 1
    U = sc["uav_num"]
    batch = "Vs" in sc
    coupled = args.workload == "circle64c"
    mode = 2 if batch else (1 if coupled else 0)
    sharded = U > 1 and world > 1 and not batch
    s = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    t0 = time.time()
    if batch:
        s.init_pointclouds(sc["Vs"])
    else:
        s.init_pointcloud(sc["V"])
    build_s = time.time() - t0
    build_ms, build_pts = s.build_stats()
    st0 = [scenes.init_state(scenes.init_spline_single(wp)) for wp in sc["way_points"]] if batch else scenes.initial_states(sc)
    first, count = 0, U
    if sharded:
        from trajopt import dist as tdist
        first, count = tdist.attach_nccl(s)          # the library's own NCCL communicator; exchanges run inside its CUDA graph
    ext = torch.cuda.ExternalStream(s.stream(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if os.environ.get("TRAJOPT_BENCH_NOFLUSH"):      # experiment only (cold- vs warm-cache kernel times); never a bench value
        flush = torch.empty(16, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident iterations: `value` (iterations W .. W+K-1 from the initial state)
    s.states_upload(st0)
    for _ in range(args.warmup):
        s.iterate(1, mode)
    s.reset_counters()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    sampler.begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gn = 0.0
    for a, b in ev:
        flush.fill_(1)                      # L2 flush, outside the timed bracket
        torch.cuda.synchronize()
        sampler.hold()
        with torch.cuda.stream(ext):
            a.record()
            gn = s.iterate(1, mode)
            b.record()
        sampler.release()
    barrier()
    sampler.end()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    ctr = s.counters()

    # ---- end to end through the host-in/host-out entry point: the SAME iterations (restart from the initial state)
    cur = [dict(x, spline=x["spline"].copy(order="F"), p_slack=x["p_slack"].copy(order="F"), t_slack=x["t_slack"].copy(),
                p_lambda=x["p_lambda"].copy(order="F"), t_lambda=x["t_lambda"].copy()) for x in st0]
    bound = s.bind_states(cur)              # tob_state array over the caller's host buffers, built once like a C++ caller would
    for _ in range(args.warmup):
        s.optimization_bound(bound, mode)
    barrier()
    sampler.begin()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                      # same L2 policy as the resident loop; not inside the timed call
        torch.cuda.synchronize()
        sampler.hold()
        t0 = time.perf_counter()
        s.optimization_bound(bound, mode)   # host buffers -> pinned staging -> H2D -> one ADMM iteration -> D2H -> host buffers
        e2e_s += time.perf_counter() - t0
        sampler.release()
    barrier()
    sampler.end()
    sampler.close()
    T = s.T
    state_bytes = U * (3 * T + 1 + 18 * P + P + 18 * P + P) * 8

    # ---- per-kernel timing pass (same problem, following iterations) for the roofline of the dominant kernel
    s.states_upload(st0)
    for _ in range(args.warmup):
        s.iterate(1, mode)
    s.profile_enable(True)
    s.reset_counters()
    psteps = min(args.steps, 10)
    for _ in range(psteps):
        flush.fill_(1)
        torch.cuda.synchronize()
        s.iterate(1, mode)
    prof = s.profile_read()
    s.profile_enable(False)
    pctr = s.counters()
    fp64_peak = s.fp64_peak_tflops()

    # max over ranks; pair counters: every rank counts the pairs of its own rows -> sum over ranks
    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    pe = torch.tensor([float(ctr["dcd_candidates"] + ctr["ccd_candidates"] + ctr["energy_plane_evals"]), float(ctr["kernel_launches"])],
                      dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pe, op=dist.ReduceOp.SUM)
    total_ms, e2e_s = float(tt[0]), float(tt[1])

    def teardown():
        # orderly: the context (and the NCCL communicator it owns) goes first, on every rank, then torch's process group
        sys.stdout.flush()
        s.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return

    # sharded: ONE problem over all ranks (strong); batch: every problem of every rank iterates once per step (strong: the
    # set of problems is fixed); else N replicas (weak)
    mult = 1 if sharded else (sc["n_total"] if batch and not args.emulate_rank else (U if batch else world))
    value = mult * args.steps / (total_ms * 1e-3)
    pair_evals = float(pe[0])      # all ranks, all problems / robots / replicas of the timed steps
    # ---- roofline: every kernel against its bound, `roofline` = the kernel with the largest share of device time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json (burst copy)" if "hbm_gbs" in peaks else "fallback of B200_PROFILING.md"
    per_step = {k: pctr[k] / float(psteps) for k in pctr if not isinstance(pctr[k], list)}
    geo = {"rows": (count if sharded else U) * P * 8, "P": P, "T": T, "U": count if sharded else U}
    models = kernel_models(per_step, geo)
    tot_prof_ms = sum(v[0] for v in prof.values())
    ncu = {}
    try:
        tkey = ("batch%d" % U) if batch else sc["name"]     # a capture is specific to the problem set on one GPU
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json"))).get(tkey, {})
    except Exception:
        pass
    ktab = {}
    for name, (kms, kn) in prof.items():
        if not kn:
            continue
        flop, byts, bound = models.get(name, (0.0, 0.0, "hbm"))
        sec = kms * 1e-3 / psteps                              # all launches of this kernel in one step
        m = ncu.get(name, {})
        dram = m.get("dram_bytes")                             # measured DRAM read+write per launch (ncu --set full)
        lps = kn / float(psteps)
        ktab[name] = {"ms_per_step": kms / psteps, "launches_per_step": lps, "bound": bound,
                      "tflops_counted": flop / sec / 1e12, "gbs_algorithmic": byts / sec / 1e9,
                      "gbs_dram_ncu": (dram * lps / sec / 1e9) if dram else None,
                      "fp64_pipe_active_pct_ncu": m.get("fp64_pipe_pct"),
                      "frac": (flop / sec / 1e12 / fp64_peak) if bound == "fp64" and fp64_peak else byts / sec / 1e9 / hbm_peak,
                      "share_of_step": kms / tot_prof_ms if tot_prof_ms else None}
        if name == "k_narrow":      # the 49-DOP gate runs through a single-precision filter: its work is not in the FP64 figure
            ktab[name]["tflops_fp32_gate_counted"] = (1.0 if GATE_FP32 else 0.0) * FLOP_KDOP_GROUP * per_step.get("np_kdop_groups", 0.0) / sec / 1e12
    roof = None
    if ktab:
        name = max(ktab, key=lambda k: ktab[k]["ms_per_step"])
        kt = ktab[name]
        lps = kt["launches_per_step"]
        fp = kt["bound"] == "fp64"
        m = ncu.get(name, {})
        roof = {"kernel": name, "bound": "fp64" if fp else "hbm",
                "achieved": kt["tflops_counted"] if fp else kt["gbs_algorithmic"], "peak": fp64_peak if fp else hbm_peak,
                "unit": "TFLOP/s" if fp else "GB/s", "frac": kt["frac"], "traffic": m.get("dram_bytes"),
                "peak_source": "FP64 DFMA microbenchmark run in this process (tob_fp64_peak); no FP64 figure in MEASURED_PEAKS.json" if fp else peak_src,
                "share_of_step": kt["share_of_step"], "avg_launch_ms": kt["ms_per_step"] / lps, "launches_per_step": lps,
                "algorithmic_flop_per_launch": models[name][0] / lps if name in models else None,
                "algorithmic_bytes_per_launch": models[name][1] / lps if name in models else None,
                "fp64_pipe_active_pct_ncu": m.get("fp64_pipe_pct"),
                "hbm": {"achieved_algorithmic": kt["gbs_algorithmic"], "achieved_dram_ncu": kt["gbs_dram_ncu"], "peak": hbm_peak,
                        "unit": "GB/s", "frac": kt["gbs_algorithmic"] / hbm_peak, "peak_source": peak_src},
                "work": "flops from device counters of the work really executed (49-DOP groups, GJK rounds, in-band barrier "
                        "terms), see kernel_models(); ncu columns come from profiles/ncu_metrics.json (captured once per round)"}
    build = None
    if build_ms and build_pts:
        bsec = build_ms * 1e-3
        build = {"points": int(build_pts), "device_ms": build_ms, "host_wall_s": build_s, "gbs_algorithmic": BYTES_PER_BUILD_POINT * build_pts / bsec / 1e9,
                 "frac_hbm": BYTES_PER_BUILD_POINT * build_pts / bsec / 1e9 / hbm_peak, "h2d_gbs": 24.0 * build_pts / bsec / 1e9,
                 "note": "one-time per problem; the device time includes the host-to-device copy of the clouds (24 B/pt), which bounds it"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if (sharded or batch) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": describe(args, world),
        "pair_evals_per_s": pair_evals / (total_ms * 1e-3),
        "pairs_per_step": {k: ctr[k] / args.steps for k in ("dcd_candidates", "planes", "ccd_candidates", "energy_plane_evals", "barrier_terms",
                                                             "np_kdop_groups", "np_gjk_iters", "np_kdop_exact", "np_band", "ccd_kdop_pass", "ccd_gjk_iters", "line_search_trials", "ls_rungs_skipped")},
        "e2e": {"value": mult * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": state_bytes * world, "d2h_bytes_per_step": state_bytes * world},
        "ls_rung_hist_per_step": [x / args.steps for x in ctr.get("ls_rung_hist", [])],
        "gpu_launches": int(pe[1]),
        "clocks": sampler.summary(),
        "roofline": roof,
        "kernels": ktab,
        "build": build,
        "run": {"problems_on_rank0": U, "points_on_rank0": int(sc.get("n_points", sc["V"].shape[0] if "V" in sc else 0)), "gnorm_last": gn,
                "emulated_rank": args.emulate_rank},
    }
    # CPU baseline on a bounded sample, rank 0, N == 1 only
    if world == 1 and not args.no_cpu:
        if batch:
            out["cpu_baseline"] = cpu_batch(sc["n_total"], warm=1, steps=2, budget_s=6.0)
        else:
            sc["coupled"] = coupled
            out["cpu_baseline"] = cpu_single(sc, P, budget_s=25.0, max_iters=6)
    OUT.emit(json.dumps(out))
    teardown()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.workload == "batch":
        cb = cpu_batch(args.problems or 1024, warm=min(args.warmup, 3), steps=args.steps, budget_s=100.0)
        steps = args.steps
    else:
        sc = workload(args, 0, 1)
        sc["coupled"] = args.workload == "circle64c"
        P = len(sc["way_points"][0]) - 1
        cb = cpu_single(sc, P, budget_s=150.0, max_iters=args.steps, warm=args.warmup)
        steps = args.steps
    value = cb["value"]
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
           "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong" if args.workload in ("batch", "circle64", "circle64c", "cross8") else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": describe(args, world),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    OUT.emit(json.dumps(out))


class QuietStdout:
    """the driver reads ONE JSON line from stdout: whatever libraries print there while the bench runs (NCCL's version banner
    at the first communicator, ...) is sent to stderr; the result line goes to the real stdout"""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.fd, (line + "\n").encode())


OUT = None


def main():
    global OUT
    OUT = QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch", choices=["batch", "forest", "bridge", "cross8", "circle64", "circle64c"])
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--problems", type=int, default=None, help="--workload batch: number of independent problems (default 1024)")
    ap.add_argument("--emulate-rank", default=None, help="--workload batch, one GPU: run the share of rank R of W ('R/W'); profiling aid")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
