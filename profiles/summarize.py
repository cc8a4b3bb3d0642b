"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.
usage: python profiles/summarize.py launches <launches.csv> <out.txt>
       python profiles/summarize.py full <file.ncu-rep> <out.txt>"""
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
