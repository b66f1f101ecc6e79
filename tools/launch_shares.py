#!/usr/bin/env python
"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_shares.py gpurun_out/launches.csv [first_id last_id]  > profiles/xyz_launch_shares.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
iK, iV, iID, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Grid Size"), hdr.index("Block Size")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows:
    if not (lo <= int(r[iID]) <= hi):
        continue
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    tot[name] = tot.get(name, 0.0) + float(r[iV].replace(",", "")) * 1e-6
    cnt[name] += 1
T = sum(tot.values())
print("launches %d (ids %d..%d), total device time %.3f ms (cold-cache, serialised: compare SHARES)" % (sum(cnt.values()), lo, min(hi, len(rows) - 1), T))
print("%-44s %6s %12s %10s %7s" % ("kernel", "n", "total ms", "ms/launch", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-44s %6d %12.3f %10.3f %6.1f%%" % (k[:44], cnt[k], v, v / cnt[k], 100 * v / T))
