#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from `ncu -i rep --page source --csv` output (file argument)."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 6]
iS, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
print("total samples", tot, " warp instructions", sum(int(r[iex]) for r in data if r[iex].isdigit()))
top = sorted([(int(r[iS]), i, r[isrc].strip()) for i, r in enumerate(data) if r[iS].isdigit()], reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for s, i, src in top:
    print("%6d %5.1f%%  #%-5d %s" % (s, 100.0 * s / max(tot, 1), i, src))
c, ci = Counter(), Counter()
for r in data:
    if r[iS].isdigit():
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[iS])
        ci[op] += int(r[iex])
print("by opcode (samples, executed):", [(k, v, ci[k]) for k, v in c.most_common(14)])
