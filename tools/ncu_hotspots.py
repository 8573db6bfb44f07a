#!/usr/bin/env python
"""Join an ncu SASS page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -gi
line info and aggregate executed instructions / stall samples per source line (innermost and
outermost inlining frame).  Usage: ncu_hotspots.py sass.csv disasm.txt kernel_substring"""
import csv
import re
import sys
from collections import defaultdict


def main(sass_csv, dis, kname):
    # address -> (inner file:line, outer line)
    addr2 = {}
    cur = None
    inside = False
    inner = outer = None
    for line in open(dis):
        if line.startswith("//---") and ".text." in line:
            inside = kname in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            f = m.group(1).split("/")[-1]
            if m.group(3):  # an inlined frame: first such line of a group is the innermost
                if cur is None:
                    cur = (f, int(m.group(2)))
            else:
                inner = cur or (f, int(m.group(2)))
                outer = (f, int(m.group(2)))
                cur = None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", line)
        if m and outer:
            addr2[int(m.group(1), 16)] = (inner, outer, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    hdr = None
    agg_in = defaultdict(lambda: [0, 0, 0])
    agg_out = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    take = False
    for r in rows:
        if r and r[0] == "Kernel Name":
            take = kname_match(r[1], kname)
            continue
        if r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if not take or hdr is None or not r or not r[0].startswith("0x") and not re.match(r"^[0-9a-f]+$", r[0]):
            continue
        a = int(r[0], 16) if not r[0].startswith("0x") else int(r[0], 16)
        base = getattr(main, "base", None)
        if base is None:
            main.base = base = a
        off = a - base
        ie = int(float(r[hdr["Instructions Executed"]] or 0))
        te = int(float(r[hdr["Thread Instructions Executed"]] or 0))
        sm = int(float(r[hdr["# Samples"]] or 0))
        info = addr2.get(off)
        if info is None:
            continue
        for agg, key in ((agg_in, info[0]), (agg_out, info[1])):
            agg[key][0] += ie; agg[key][1] += te; agg[key][2] += sm
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print(f"total warp-instr {tot[0]:,}  thread-instr {tot[1]:,}  samples {tot[2]:,}  avg active {tot[1]/max(tot[0],1):.1f}")
    for name, agg in (("OUTER (kernel body line)", agg_out), ("INNER (innermost inlined line)", agg_in)):
        print("==", name)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
            print(f"  {k[0]}:{k[1]:<5d} instr {100*v[0]/tot[0]:5.1f}%  samples {100*v[2]/max(tot[2],1):5.1f}%  active {v[1]/max(v[0],1):4.1f}")


def kname_match(full, sub):
    return True


if __name__ == "__main__":
    main(*sys.argv[1:4])
