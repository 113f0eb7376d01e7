#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the number
of launches, total and share of device time.  The torch generator kernels (synthetic input) are
listed separately from the xsi:: kernels of the path.  usage: tools/launches_summary.py launches.csv"""
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = r[4].replace("void ", "")
        name = re.sub(r"\((xsi::|unsigned|int|const).*", "", name).replace("(int)", "")
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += float(r[14]) / 1e6
    pat = re.compile(r"^(xsi::)?(scan_rows|build_wah|pbwt_|wah_|scan_u32|sparse_|pack_wah|compose_)")
    ours = {k: v for k, v in agg.items() if pat.match(k)}  # newer ncu prints the names without the namespace
    tot = sum(v[1] for v in ours.values())
    print("# xsi:: kernels: %d launches, %.3f ms (ncu-serialised, cold cache); other (torch generator / fills): %d launches, %.3f ms"
          % (sum(v[0] for v in ours.values()), tot, sum(v[0] for k, v in agg.items() if k not in ours),
             sum(v[1] for k, v in agg.items() if k not in ours)))
    print("%-44s %8s %12s %7s  %s" % ("kernel", "launches", "total ms", "share", "block x grid (last)"))
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.3f %6.1f%%  %s x %s" % (k, v[0], v[1], 100 * v[1] / tot, v[2], v[3]))


if __name__ == "__main__":
    main()
