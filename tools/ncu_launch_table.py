#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: ncu_launch_table.py launches.csv [first_timed_launch_id]  (> profiles/rNN_launches.txt)
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py, not absolutes."""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
tot = 0.0
for r in rows[1:]:
    if int(r[ix["ID"]]) < skip:
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    ns = float(r[ix["Metric Value"]])
    a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
    a[0] += 1
    a[1] += ns
    tot += ns
print("%-44s %8s %12s %10s %7s  %s" % ("kernel", "launches", "total us", "avg us", "share", "grid x block"))
for n, (c, ns, g, b) in agg.items():
    print("%-44s %8d %12.1f %10.2f %6.1f%%  %s x %s" % (n, c, ns / 1e3, ns / 1e3 / c, 100 * ns / tot, g, b))
print("%-44s %8d %12.1f" % ("total", sum(a[0] for a in agg.values()), tot / 1e3))
