#!/usr/bin/env python
"""Condenses ncu outputs brought back in gpurun_out/ into small text summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-72s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %6d %12.1f %10.1f %7.3f" % (n, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[1] / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    print("# ncu --set full --clock-control none : per-launch raw metrics")
    for r in rows[2:]:
        print("== " + r[ki][:90])
        for w, i in idx:
            print("   %-78s %16s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
