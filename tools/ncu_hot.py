#!/usr/bin/env python
"""Top SASS instructions of one kernel by warp-stall samples, from an .ncu-rep (read here, no GPU).
usage: tools/ncu_hot.py report.ncu-rep kernel-regex [top-n]"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # first kernel only
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[hdr_i[0]]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    body = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
    si = h.index("# Samples")
    ii = h.index("Instructions Executed")
    stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si]) for r in body)
    toti = sum(int(r[ii]) for r in body)
    print("kernel:", rows[hdr_i[0] - 1][1] if hdr_i[0] else "?", "| samples", tot, "| warp instr", toti)
    agg = {}
    for r in body:
        for i in stall:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, tot)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][si]))[:top]:
        st = sorted(((int(r[i]), h[i][6:]) for i in stall), reverse=True)[:2]
        print("%5d %5.1f%% inst %11s  %-58s %s" % (idx, 100.0 * int(r[si]) / max(1, tot), r[ii], r[1].strip()[:58],
                                                " ".join("%s:%d" % (n, v) for v, n in st if v)))


if __name__ == "__main__":
    main()
