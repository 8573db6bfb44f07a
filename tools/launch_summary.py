#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list per kernel (time share, instructions,
active lanes, issue utilisation).  Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for x in csv.DictReader(lines):
        key = (x["ID"], re.sub(r"\(.*", "", x["Kernel Name"])[:60])
        launches.setdefault(key, {})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for (_, k), m in launches.items():
        a = agg[k]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0) / 1e6
        a[2] += m.get("smsp__inst_executed.sum", 0)
        a[3] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0)
        a[4] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} n={a[0]:3d} ms={a[1]:8.3f} ({100 * a[1] / tot:4.1f}%) avg={a[1] / a[0]:.3f} "
              f"inst={a[2] / a[0]:.3e} lanes={a[3] / a[0]:.1f} issue={a[4] / a[0]:.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
