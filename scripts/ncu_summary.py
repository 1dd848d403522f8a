#!/usr/bin/env python
"""Summarise ncu outputs brought back from a gpurun visit.

  python scripts/ncu_summary.py launches <launches.csv>      # per-kernel share of the step
  python scripts/ncu_summary.py raw <prof.ncu-rep>           # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000.0
        elif r[ui] == "ms":
            v *= 1000.0
        agg[r[ki].split("(")[0][:100]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'share':>7} {'n':>5} {'avg us':>10} {'max us':>10}  kernel   (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) / tot * 100:6.2f}% {len(v):5d} {sum(v) / len(v):10.1f} {max(v):10.1f}  {k}")


def raw(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[ni])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:82s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
