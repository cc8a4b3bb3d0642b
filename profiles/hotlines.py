"""usage: python profiles/hotlines.py <file.ncu-rep> <kernel-name> [top]
Top source lines of one kernel by warp-stall samples (ncu --import-source on; kernels built with -lineinfo)."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    agg = {}
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if r and r[0] == "Line No":
            hdr = r
            i_s = hdr.index("# Samples")
            i_i = hdr.index("Instructions Executed")
            continue
        if hdr is None or not r or r[0] == "" or not r[0].isdigit():
            continue
        try:
            key = (fname, int(r[0]), r[1].strip()[:110])
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[i_s]); a[1] += int(r[i_i])
        except (ValueError, IndexError):
            pass
    tot = sum(v[0] for v in agg.values()) or 1
    print("# %s : %d samples" % (kern, tot))
    for (f, ln, src), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% %8d inst  %s:%d  %s" % (100.0 * s / tot, n, f, ln, src))


main()
