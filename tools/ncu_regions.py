#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump (SASS view) by address ranges.
usage: ncu_regions.py src.csv [off0:off1:name ...]   (hex offsets relative to the first instruction)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
base = int(body[0][0], 16)
regions = []
for a in sys.argv[2:]:
    lo, hi, name = a.split(":")
    regions.append((int(lo, 16), int(hi, 16), name))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[ix["# Samples"]]) for r in body)
print("total warp-instructions %d, samples %d" % (tot_inst, tot_samp))
for lo, hi, name in regions:
    inst = samp = thr = 0
    st = {c: 0 for c in stall_cols}
    n = 0
    for r in body:
        off = int(r[0], 16) - base
        if lo <= off < hi:
            n += 1
            inst += int(r[ix["Instructions Executed"]])
            thr += int(r[ix["Thread Instructions Executed"]])
            samp += int(r[ix["# Samples"]])
            for c in stall_cols:
                st[c] += int(r[ix[c]])
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print("%-14s sass=%4d inst=%10d (%5.1f%%) avg_thr=%5.1f samples=%7d (%5.1f%%)  %s" % (
        name, n, inst, 100.0 * inst / tot_inst, thr / max(1, inst), samp, 100.0 * samp / tot_samp,
        " ".join("%s=%d" % (k[6:], v) for k, v in top)))
