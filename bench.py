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

# algorithmic FP64 work per unit
FLOP_PER_DCD_CANDIDATE = 3.0e3    # k-DOP 49x(7x5+4) worst case + GJK(6,1) ~3-6 iterations
BYTES_PER_DCD_CANDIDATE = 28 + 32  # point + id read, plane write when accepted
FLOP_PER_CCD_CANDIDATE = 3.4e3 + 1.5e3  # swept k-DOP on 12 points + >= one GJK(12,1)


def workload(name, n_pts=None):
    from trajopt import scenes
    if name == "forest":
        sc = scenes.forest(n_pts=n_pts or 1_000_000)
    elif name == "bridge":
        sc = scenes.bridge(n_pts=n_pts or 100_000)
    elif name == "circle64":      # BASELINE.json configs[3]: 64 UAVs, inter-robot planes, robots sharded over the ranks
        sc = scenes.circle(n_uav=64, n_pts=n_pts or 20_000)
    elif name == "cross8":        # configs[2]
        sc = scenes.cross(n_pts=n_pts or 50_000)
    else:
        raise SystemExit("unknown workload " + name)
    return sc


def clock_sampler(stop, out):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200", "-i",
                              os.environ.get("LOCAL_RANK", "0")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()


def summarize_clocks(lines):
    sm, smax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1])); smax.append(float(f[2]))
        except ValueError:
            continue
        for i, nm in enumerate(names):
            if f[5 + i].lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


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

    sc = workload(args.workload, args.points)
    P = len(sc["way_points"][0]) - 1
    U = sc["uav_num"]
    sharded = U > 1 and world > 1
    s = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    t0 = time.time()
    s.init_pointcloud(sc["V"])
    build_s = time.time() - t0
    st0 = scenes.initial_states(sc)
    if sharded:
        from trajopt import dist as tdist
        tdist.attach(s)
    ext = torch.cuda.ExternalStream(s.stream(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident iterations: `value`
    s.states_upload(st0)
    for _ in range(args.warmup):
        s.iterate(1)
    s.reset_counters()
    lines, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, lines), daemon=True)
    th.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gn = 0.0
    for a, b in ev:
        flush.fill_(1)                      # L2 flush, outside the timed bracket
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            a.record()
            gn = s.iterate(1)
            b.record()
    barrier()
    stop.set()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    ctr = s.counters()

    # ---- per-kernel timing pass (same problem, next iterations) for the roofline of the dominant kernel
    s.profile_enable(True)
    s.reset_counters()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        s.iterate(1)
    prof = s.profile_read()
    s.profile_enable(False)
    pctr = s.counters()
    fp64_peak = s.fp64_peak_tflops()

    # ---- end to end through the host-in/host-out entry point
    cur = [pinned_state(x) for x in s.states_download(st0)]
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                      # same L2 policy as the resident loop; not inside the timed call
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = s.optimization(cur)           # host buffers in -> H2D -> one ADMM iteration -> D2H -> host buffers out
        e2e_s += time.perf_counter() - t0
        for c_, r_ in zip(cur, res):
            for k in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
                c_[k][...] = r_[k]
            c_["piece_time"] = r_["piece_time"]
    T = s.T
    state_bytes = U * (3 * T + 1 + 18 * P + P + 18 * P + P) * 8

    # max over ranks
    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    mult = 1 if sharded else world         # sharded: ONE problem over all ranks (strong); else N replicas (weak)
    value = mult * args.steps / (total_ms * 1e-3)
    pair_evals = ctr["dcd_candidates"] + ctr["ccd_candidates"] + ctr["energy_plane_evals"]
    # dominant kernel by device time
    dom = max(prof.items(), key=lambda kv: kv[1][0])
    tot_prof_ms = sum(v[0] for v in prof.values())
    roof = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    name, (kms, kn) = dom
    if kn:
        per_launch_s = kms * 1e-3 / kn
        if name == "k_narrow":
            units = pctr["dcd_candidates"] / kn
            flops, byts = FLOP_PER_DCD_CANDIDATE * units, BYTES_PER_DCD_CANDIDATE * units
        elif name == "k_ccd":
            units = pctr["ccd_candidates"] / kn
            flops, byts = FLOP_PER_CCD_CANDIDATE * units, 28 * units
        else:
            units, flops, byts = 0, 0.0, 0.0
        roof = {"kernel": name, "bound": "fp64", "achieved": flops / per_launch_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (flops / per_launch_s / 1e12) / fp64_peak if fp64_peak else None, "traffic": None,
                "peak_source": "DFMA microbenchmark run in this process (tob_fp64_peak)",
                "share_of_step": kms / tot_prof_ms if tot_prof_ms else None, "units_per_launch": units,
                "avg_launch_ms": per_launch_s * 1e3,
                "hbm": {"achieved": byts / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / per_launch_s / 1e9 / hbm_peak,
                        "peak_source": peak_src}}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %d UAV, %d pts, %d Bezier pieces (%d sub-segments each), 3D.json params" % (sc["name"], U, sc["V"].shape[0], P, P * 8),
                   "l2": "flushed between timed iterations (256 MiB rewrite outside the event brackets)",
                   "multi_gpu": ("robots sharded over ranks, NCCL all-gather of control points/directions" if sharded else
                                 ("replicas only" if world > 1 else "single")), "lbvh_build_s": build_s, "gnorm_last": gn},
        "pair_evals_per_s": pair_evals * mult / (total_ms * 1e-3),
        "pairs_per_step": {k: ctr[k] / args.steps for k in ("dcd_candidates", "planes", "ccd_candidates", "energy_plane_evals")},
        "e2e": {"value": mult * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes},
        "gpu_launches": int(ctr["kernel_launches"]),
        "clocks": summarize_clocks(lines),
        "roofline": roof,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]},
    }
    # CPU baseline on a bounded sample, rank 0, N == 1 only
    if world == 1 and not args.no_cpu:
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="forest")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
