#!/usr/bin/env python
"""bench.py -- ADMM iterations/s (and segment-point pair evaluations/s) of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload forest|bridge]

One "step" = one ADMM iteration (Optimization3D_admm::optimization, Optimization3D_admm.h:29-67) on the
BASELINE.json configs[1] workload: single UAV, synthetic dense-forest cloud of 1 M points, 64 Bezier pieces
(512 sub-segments), FP64, Config File/3D.json parameters, straight-line initial trajectory.

  value  : iterations/s with the state resident in HBM (tob_admm_iterate), CUDA events on the library's stream,
           L2 flushed between iterations (a 256 MiB buffer is rewritten outside the timed brackets).
  e2e    : the same metric through the reference-shaped entry point (host buffers in, host buffers out:
           tob_optimization = upload + iterate + download per call), pinned host memory, wall clock around the call.
  N > 1  : single-UAV problems do not shard ("replicas only"): every rank runs its own replica of the
           problem; value = total iterations of all ranks / max-over-ranks time ("weak").
  --impl reference : the reference's own CPU implementation (oracle/_ref = the unmodified sources compiled here,
           else the C port) on the same scene, 1 core (the reference has no parallel region), rank 0 only.
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

# ---- algorithmic work per unit; `roofline.achieved` is computed from these
FLOP_PER_DCD_CANDIDATE = 3.0e3          # 49-DOP worst case 49x(7x5+4) + GJK(6,1), ~3-6 iterations
BYTES_PER_DCD_CANDIDATE = 28 + 32       # point + id read, plane write when accepted
FLOP_PER_CCD_CANDIDATE = 3.4e3 + 1.5e3  # swept 49-DOP on 12 points + >= one GJK(12,1)
FLOP_PER_PLANE_EVAL = 36.0              # 6 control points x (3 mul + 3 add)
FLOP_PER_ACTIVE_TERM_E = 45.0           # energy: 2 sub, 3 mul, 1 div, log ~35 DFMA-equivalents
FLOP_PER_ACTIVE_TERM_G = 116.0          # gradient: e1,e2 (~95) + 3 + 6 accumulates x 2
BYTES_PER_PLANE = 32.0


def kernel_models(per_step, geo):
    """kernel name -> (algorithmic FP64 flop per step, algorithmic bytes per step, bound) for the kernels of one ADMM
    iteration.  per_step: counters of the profiled pass divided by its step count; geo: rows, n1 (level-1 nodes), P, T, U."""
    rows, n1, P, T, U = geo["rows"], geo["n1"], geo["P"], geo["T"], geo["U"]
    cand, ccd, planes = per_step["dcd_candidates"], per_step["ccd_candidates"], per_step["planes"]
    evals, terms = per_step["energy_plane_evals"], per_step["barrier_terms"]
    e_evals = max(evals - planes, 0.0)                      # line-search passes (the gradient pass streams each plane once)
    g_share = planes / evals if evals else 0.0
    n_sys = 3 * (T - 4) + 1
    bp_bytes = 48.0 * rows + 48.0 * rows * n1 + 28.0 * cand  # row box + one level-1 box per (row, node) + candidate out
    return {
        "k_rows": (0.0, rows * (18 + 6 + 2 * 49) * 8.0 * 2, "hbm"),
        "k_bp_count": (0.0, bp_bytes - 4.0 * cand, "hbm"),
        "k_bp_fill": (0.0, bp_bytes, "hbm"),
        "k_bp_ccd": (FLOP_PER_CCD_CANDIDATE * ccd, 48.0 * rows + 48.0 * rows * n1 + 24.0 * ccd, "hbm"),
        "k_narrow": (FLOP_PER_DCD_CANDIDATE * cand, BYTES_PER_DCD_CANDIDATE * cand, "fp64"),
        "k_pack": (0.0, 4.0 * cand + 72.0 * planes, "hbm"),
        "k_row_energy": (FLOP_PER_PLANE_EVAL * e_evals + FLOP_PER_ACTIVE_TERM_E * terms * (1 - g_share), BYTES_PER_PLANE * e_evals, "fp64"),
        "k_row_grad": (FLOP_PER_PLANE_EVAL * planes + FLOP_PER_ACTIVE_TERM_G * terms * g_share, BYTES_PER_PLANE * planes, "fp64"),
        "k_piece": (6.0 * (18 + 171) * 2 * rows + U * P * 19 ** 3 / 3.0, U * P * (361 + 19) * 8.0, "fp64"),
        "k_solve_bcr": (U * (n_sys * (17 ** 2 + 17) + 2 * n_sys * 17), U * P * (361 + 19) * 8.0, "fp64"),
        "k_slack": (U * P * (19 ** 3 / 3.0 + 4 * 19 * 19), U * P * (4 * 19) * 8.0, "fp64"),
        "k_robot_ls": (0.0, 9.0 * rows * 20, "hbm"),
    }


def workload(name, n_pts=None, n_problems=None):
    from trajopt import scenes
    if name == "forest":
        sc = scenes.forest(n_pts=n_pts or 1_000_000)
    elif name == "bridge":
        sc = scenes.bridge(n_pts=n_pts or 100_000)
    elif name == "circle64":      # BASELINE.json configs[3]: 64 UAVs, inter-robot planes, robots sharded over the ranks
        sc = scenes.circle(n_uav=64, n_pts=n_pts or 20_000)
    elif name == "cross8":        # configs[2]
        sc = scenes.cross(n_pts=n_pts or 50_000)
    elif name == "batch":         # configs[4]: independent single-UAV problems, clouds log-uniform in [1e4, 1e6] points
        rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
        total = n_problems or 1024
        mine = scenes.batch_partition(total, world, rank)   # balanced deal over the ranks, no communication
        ms = [scenes.batch_member(k) for k in mine]
        sc = dict(name="batch", Vs=[m["V"] for m in ms], way_points=[m["way_points"][0] for m in ms], uav_num=len(ms), ks=1e-8,
                  V=np.zeros((sum(m["V"].shape[0] for m in ms), 0)), n_total=total)
    else:
        raise SystemExit("unknown workload " + name)
    return sc


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


def pinned_state(st):
    import torch
    out = {}
    for k, v in st.items():
        if isinstance(v, np.ndarray):
            t = torch.empty(v.size, dtype=torch.float64).pin_memory()
            a = t.numpy().reshape(v.shape, order="F")
            a[...] = v
            out[k] = a
            out["_keep_" + k] = t
        else:
            out[k] = v
    return out


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

    sc = workload(args.workload, args.points, args.problems)
    P = len(sc["way_points"][0]) - 1
    U = sc["uav_num"]
    batch = "Vs" in sc
    mode = 2 if batch else 0
    sharded = U > 1 and world > 1 and not batch
    s = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    t0 = time.time()
    if batch:
        s.init_pointclouds(sc["Vs"])
    else:
        s.init_pointcloud(sc["V"])
    build_s = time.time() - t0
    st0 = [scenes.init_state(scenes.init_spline_single(wp)) for wp in sc["way_points"]] if batch else scenes.initial_states(sc)
    if sharded:
        from trajopt import dist as tdist
        tdist.attach(s)
    ext = torch.cuda.ExternalStream(s.stream(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if os.environ.get("TRAJOPT_BENCH_NOFLUSH"):      # experiment only (cold- vs warm-cache kernel times); never a bench value
        flush = torch.empty(16, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident iterations: `value`
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

    # ---- per-kernel timing pass (same problem, next iterations) for the roofline of the dominant kernel
    s.profile_enable(True)
    s.reset_counters()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        s.iterate(1, mode)
    prof = s.profile_read()
    s.profile_enable(False)
    pctr = s.counters()
    fp64_peak = s.fp64_peak_tflops()

    # ---- end to end through the host-in/host-out entry point
    cur = [pinned_state(x) for x in s.states_download(st0)]
    bound = s.bind_states(cur)              # tob_state array over the pinned host buffers, built once like a C++ caller would
    barrier()
    sampler.begin()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                      # same L2 policy as the resident loop; not inside the timed call
        torch.cuda.synchronize()
        sampler.hold()
        t0 = time.perf_counter()
        s.optimization_bound(bound, mode)   # host buffers in -> H2D -> one ADMM iteration -> D2H -> same host buffers
        e2e_s += time.perf_counter() - t0
        sampler.release()
    sampler.end()
    sampler.close()
    T = s.T
    state_bytes = U * (3 * T + 1 + 18 * P + P + 18 * P + P) * 8

    # max over ranks; pair counters: every rank counts the pairs of its own rows -> sum over ranks
    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    pe = torch.tensor([float(ctr["dcd_candidates"] + ctr["ccd_candidates"] + ctr["energy_plane_evals"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pe, op=dist.ReduceOp.SUM)
    total_ms, e2e_s = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # sharded: ONE problem over all ranks (strong); batch: every problem of every rank iterates once per step (strong: the
    # set of problems is fixed); else N replicas (weak)
    mult = 1 if sharded else (sc["n_total"] if batch else world)
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
    per_step = {k: pctr[k] / float(args.steps) for k in pctr}
    n0 = -(-sc["V"].shape[0] // 32)
    n1 = float(np.mean([-(-v.shape[0] // 1024) for v in sc["Vs"]])) if batch else -(-n0 // 32)
    geo = {"rows": U * P * 8, "n1": n1, "P": P, "T": T, "U": U}
    models = kernel_models(per_step, geo)
    tot_prof_ms = sum(v[0] for v in prof.values())
    traffic = {}
    try:
        tkey = "batch%d" % sc["n_total"] if batch else sc["name"]     # the capture is specific to the problem count
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(tkey, {})
    except Exception:
        pass
    ktab = {}
    for name, (kms, kn) in prof.items():
        if not kn:
            continue
        flop, byts, bound = models.get(name, (0.0, 0.0, "hbm"))
        sec = kms * 1e-3 / args.steps                         # all launches of this kernel in one step
        ktab[name] = {"ms_per_step": kms / args.steps, "launches_per_step": kn / float(args.steps), "bound": bound,
                      "tflops": flop / sec / 1e12, "gbs": byts / sec / 1e9,
                      "frac": (flop / sec / 1e12 / fp64_peak) if bound == "fp64" and fp64_peak else byts / sec / 1e9 / hbm_peak,
                      "share_of_step": kms / tot_prof_ms if tot_prof_ms else None}
    roof = None
    if ktab:
        name = max(ktab, key=lambda k: ktab[k]["ms_per_step"])
        kt = ktab[name]
        lps = kt["launches_per_step"]
        fp = kt["bound"] == "fp64"
        roof = {"kernel": name, "bound": "fp64" if fp else "hbm",
                "achieved": kt["tflops"] if fp else kt["gbs"], "peak": fp64_peak if fp else hbm_peak, "unit": "TFLOP/s" if fp else "GB/s",
                "frac": kt["frac"], "traffic": traffic.get(name),
                "peak_source": "FP64 DFMA microbenchmark run in this process (tob_fp64_peak); no FP64 figure in MEASURED_PEAKS.json" if fp else peak_src,
                "share_of_step": kt["share_of_step"], "avg_launch_ms": kt["ms_per_step"] / lps, "launches_per_step": lps,
                "algorithmic_flop_per_launch": models[name][0] / lps if name in models else None,
                "algorithmic_bytes_per_launch": models[name][1] / lps if name in models else None,
                "hbm": {"achieved": kt["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": kt["gbs"] / hbm_peak, "peak_source": peak_src},
                "note": "single-problem kernels of ~10-100 us: latency-bound, see DESIGN.md section 3"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if (sharded or batch) else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": ("batch: %d independent single-UAV problems (%d on this rank), clouds 1e4..1e6 pts (%d pts on this rank), %d Bezier pieces each, 3D.json params"
                                % (sc["n_total"], U, sc["V"].shape[0], P)) if batch else
                               "%s: %d UAV, %d pts, %d Bezier pieces (%d sub-segments each), 3D.json params" % (sc["name"], U, sc["V"].shape[0], P, P * 8),
                   "l2": "flushed between timed iterations (256 MiB rewrite outside the event brackets)",
                   "multi_gpu": ("robots sharded over ranks, NCCL all-gather of control points/directions" if sharded else
                                 ("independent problems dealt round-robin to the ranks, no communication" if batch else
                                  ("replicas only" if world > 1 else "single"))), "lbvh_build_s": build_s, "gnorm_last": gn},
        "pair_evals_per_s": pair_evals / (total_ms * 1e-3),
        "pairs_per_step": {k: ctr[k] / args.steps for k in ("dcd_candidates", "planes", "ccd_candidates", "energy_plane_evals", "barrier_terms")},
        "e2e": {"value": mult * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes},
        "gpu_launches": int(ctr["kernel_launches"]),
        "clocks": sampler.summary(),
        "roofline": roof,
        "kernels": ktab,
    }
    # CPU baseline on a bounded sample, rank 0, N == 1 only
    if world == 1 and not args.no_cpu and not batch:
        out["cpu_baseline"] = cpu_baseline(sc, P, budget_s=25.0, max_iters=6)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(sc, P, budget_s, max_iters, sample_points=None):
    """the reference's CPU path timed on this box's host cores (1 core: the reference has no parallel region)"""
    from oracle import oracle_api as oa
    from trajopt import scenes
    o = oa.get()
    o.setup(oa.Params(P, uav_num=sc["uav_num"], ks=sc["ks"]))
    V = sc["V"]
    # the reference's incremental tree build is O(minutes) for 1 M points in random order; Morton-free trick is not
    # available to it, so the build (one-time, outside the metric) is done on the full cloud and not timed.
    t0 = time.time()
    o.init_pointcloud(V)
    build = time.time() - t0
    sts = scenes.initial_states(sc)
    n, t_used = 0, 0.0
    while n < max_iters and t_used < budget_s:
        t0 = time.perf_counter()
        sts = o.optimization_multi(sts, coupled=False) if len(sts) > 1 else [o.optimization(sts[0])]
        t_used += time.perf_counter() - t0
        n += 1
    return {"value": n / t_used, "unit": UNIT, "cores": 1, "kind": o.kind,
            "sample": "first %d ADMM iterations of the same scene from the same initial state (tree build %.1f s excluded)" % (n, build)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = workload(args.workload, args.points)
    P = len(sc["way_points"][0]) - 1
    from oracle import oracle_api as oa
    from trajopt import scenes
    o = oa.get()
    o.setup(oa.Params(P, uav_num=sc["uav_num"], ks=sc["ks"]))
    t0 = time.time()
    o.init_pointcloud(sc["V"])
    build = time.time() - t0
    sts = scenes.initial_states(sc)
    step = (lambda x: o.optimization_multi(x, coupled=False)) if len(sts) > 1 else (lambda x: [o.optimization(x[0])])
    budget = 150.0
    t_all = 0.0
    for _ in range(args.warmup):
        t0 = time.perf_counter(); sts = step(sts); t_all += time.perf_counter() - t0
        if t_all > budget / 3:
            break
    n, t_used = 0, 0.0
    while n < args.steps and t_used < budget:
        t0 = time.perf_counter()
        sts = step(sts)
        t_used += time.perf_counter() - t0
        n += 1
    value = n / t_used
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = "%d ADMM iterations (time-bounded from --steps %d) on 1 host core, tree build %.1f s excluded" % (n, args.steps, build)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": n, "warmup": args.warmup,
           "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s: %d UAV, %d pts, %d Bezier pieces (%d sub-segments each), 3D.json params" % (sc["name"], sc["uav_num"], sc["V"].shape[0], P, P * 8)},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": o.kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="forest")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--problems", type=int, default=None, help="--workload batch: number of independent problems (default 1024)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
