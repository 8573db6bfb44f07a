#!/usr/bin/env python
"""Decode the scheduling control fields of sm_100 SASS (cuobjdump -sass output): stall count, yield, write/read
dependency barrier set by the instruction, barrier wait mask.  Usage: sass_ctrl.py file.sass [lo hi] (hex addresses)"""
import re, sys
L = open(sys.argv[1]).read().splitlines()
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
for k, l in enumerate(L):
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', l)
    if not m: continue
    a = int(m.group(1), 16)
    if a < lo or a > hi: continue
    h = int(re.search(r'/\* (0x[0-9a-f]{16}) \*/', L[k + 1]).group(1), 16)
    c = (h >> 41) & 0x7fffff
    wb, rb, wait = (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3f
    print("%04x st=%2d y=%d W=%s R=%s wait=%s  %s" % (a, c & 0xf, (c >> 4) & 1, wb if wb != 7 else '-', rb if rb != 7 else '-',
          ''.join(str(b) for b in range(6) if wait >> b & 1) or '-', m.group(2).strip()[:100]))
