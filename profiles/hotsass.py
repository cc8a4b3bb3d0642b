"""usage: python profiles/hotsass.py <file.ncu-rep> <kernel-name> [top]
Top SASS instructions of one kernel by warp-stall samples, with the source line they belong to (ncu --import-source on,
-lineinfo) and every per-instruction column of the source page that is not zero (stall reasons, executed counts)."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    body = []
    for r in rows:
        if r and "# Samples" in r and hdr is None:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            body.append(r)
    if not hdr:
        print("no source page for", kern); print(out[:2000]); return
    i_s = hdr.index("# Samples")
    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    tot = sum(num(r[i_s]) for r in body) or 1.0
    print("# %s: %d SASS instructions, %d samples; columns: %s" % (kern, len(body), tot, " | ".join(hdr)))
    # opcode histogram by samples
    agg = {}
    i_src = hdr.index("Source") if "Source" in hdr else 1
    for r in body:
        op = r[i_src].strip().split()
        op = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")).split(".")[0]
        a = agg.setdefault(op, [0.0, 0])
        a[0] += num(r[i_s]); a[1] += 1
    print("# samples by opcode:", ", ".join("%s %.1f%%" % (k, 100 * v[0] / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]))
    for r in sorted(body, key=lambda r: -num(r[i_s]))[:top]:
        extra = ["%s=%s" % (hdr[i][:28], r[i]) for i in range(len(hdr)) if i not in (i_s, i_src) and r[i] not in ("", "0", "0.0") and num(r[i]) != 0.0][:10]
        print("%5.2f%%  %-70s %s" % (100 * num(r[i_s]) / tot, r[i_src].strip()[:70], " ".join(extra)))


main()
