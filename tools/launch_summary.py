#!/usr/bin/env python
"""Aggregate an ncu launch list (ncu --metrics gpu__time_duration.sum[,...] --csv --log-file X.csv) per kernel: launches, mean
duration, share of the summed device time, and -- when captured -- warp instructions and active lanes per instruction."""
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        k = r[ix["Kernel Name"]]
        k = k[:k.index("(")] if "(" in k else k
        v = float(r[ix["Metric Value"]].replace(",", ""))
        agg.setdefault(k, {}).setdefault(r[ix["Metric Name"]], []).append(v)
    tot = sum(sum(v.get("gpu__time_duration.sum", [0])) for v in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'mean ms':>8s} {'share':>6s} {'warp instr':>11s} {'lanes':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        t = v.get("gpu__time_duration.sum", [0])
        ins = v.get("smsp__inst_executed.sum")
        lanes = v.get("smsp__thread_inst_executed_per_inst_executed.ratio")
        print(f"{k[:72]:72s} {len(t):4d} {sum(t) / len(t) / 1e6:8.3f} {100 * sum(t) / tot:5.1f}% "
              f"{(sum(ins) / len(ins)) if ins else float('nan'):11.3g} {(sum(lanes) / len(lanes)) if lanes else float('nan'):6.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
