"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.
usage: python profiles/summarize.py launches <launches.csv> <out.txt>
       python profiles/summarize.py full <file.ncu-rep> <out.txt>
       python profiles/summarize.py rawcsv <raw.csv from `ncu -i rep --page raw --csv`> <out.txt> [<workload name>]
           (with a workload name: also folds the per-launch DRAM traffic of every kernel into profiles/ncu_traffic.json,
            which bench.py reads for `roofline.traffic`)"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "lts__t_bytes.sum"]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        name = r[ik].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; total %.1f us over %d launches\n" % (src, tot / 1e3, sum(v[0] for v in agg.values())))
        f.write("%-60s %8s %12s %8s %10s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %7.1f%% %10.2f\n" % (k[:60], n, t / 1e3, 100 * t / tot, t / 1e3 / n))


def full(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("kernel: %s\n" % d.get("Kernel Name"))
            for k in WANT:
                if k in d:
                    f.write("  %-70s %s %s\n" % (k, d[k], rows[1][hdr.index(k)]))
            for k in hdr:
                if "issue_stalled" in k and "per_issue_active" in k:
                    try:
                        if float(d[k]) > 0.25:
                            f.write("  stall %-64s %s\n" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[k]))
                    except ValueError:
                        pass
            f.write("\n")


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def _kname(full):
    """'void tob::k_narrow<4, true>(NarrowArgs)' -> 'k_narrow' (the names bench.py uses)"""
    n = full.split("(")[0].replace("void ", "").replace("tob::", "").strip()
    return n.split("<")[0]


def rawcsv(src, dst, workload=None):
    import json
    import os
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    per = defaultdict(list)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        per[_kname(d.get("Kernel Name", "?"))].append(d)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none; source %s ; mean over the captured launches of each kernel\n" % src)
        for k, lst in sorted(per.items()):
            f.write("kernel: %s  (%d launches captured)\n" % (k, len(lst)))
            for m in WANT:
                if m in hdr:
                    vals = [_num(d[m]) for d in lst if _num(d[m]) is not None]
                    if vals:
                        f.write("  %-66s %14.6g %s\n" % (m, sum(vals) / len(vals), units[hdr.index(m)]))
            for m in hdr:
                if "issue_stalled" in m and "per_issue_active" in m:
                    vals = [_num(d[m]) for d in lst if _num(d[m]) is not None]
                    if vals and sum(vals) / len(vals) > 0.4:
                        f.write("  stall %-60s %.2f\n" % (m.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), sum(vals) / len(vals)))
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if m in hdr:
                    u = scale.get(units[hdr.index(m)], 1.0)
                    vals = [_num(d[m]) * u for d in lst if _num(d[m]) is not None]
                    tot += sum(vals) / max(len(vals), 1)
            traffic[k.replace("tob::", "")] = tot
            f.write("  dram bytes per launch (read+write)                                   %.0f\n\n" % tot)
    if workload:
        # per-kernel measured columns for bench.py (`kernels[*].gbs_dram_ncu`, `fp64_pipe_active_pct_ncu`, `roofline.traffic`)
        mj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_metrics.json")
        cur = json.load(open(mj)) if os.path.exists(mj) else {}
        ent = {}
        for k, lst in per.items():
            def mean(m):
                vals = [_num(d[m]) for d in lst if m in d and _num(d[m]) is not None]
                return sum(vals) / len(vals) if vals else None
            ent[k.replace("tob::", "")] = {"dram_bytes": traffic[k.replace("tob::", "")],
                                           "fp64_pipe_pct": mean("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                           "warps_active_pct": mean("sm__warps_active.avg.pct_of_peak_sustained_active"),
                                           "lanes_per_inst": mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
                                           "registers": mean("launch__registers_per_thread"), "launches_captured": len(lst)}
        cur[workload] = ent
        cur["_note"] = ("per launch, from one `ncu --set full --clock-control none` capture per round (profiles/run_ncu_r02.sh): "
                        "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum; keyed by bench workload (batch<problems on the GPU>), then kernel")
        json.dump(cur, open(mj, "w"), indent=1, sort_keys=True)
        tj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
        cur = json.load(open(tj)) if os.path.exists(tj) else {}
        cur[workload] = traffic
        cur.setdefault("_note", "dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture "
                                "(profiles/run_ncu.sh); keyed by bench workload name, then kernel")
        json.dump(cur, open(tj, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"launches": launches, "full": full, "rawcsv": rawcsv}[sys.argv[1]](*sys.argv[2:])
