#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one line per profiled launch with the counters
DESIGN.md quotes.  usage: tools/ncu_summary.py report.ncu-rep [more metric names]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum"]


def main():
    rep = sys.argv[1]
    want = WANT + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for row in rows[2:]:
        print("==", row[h.index("Kernel Name")][:70])
        for w in want:
            if w in h:
                i = h.index(w)
                print("   %-62s %18s %s" % (w, row[i], units[i]))


if __name__ == "__main__":
    main()
