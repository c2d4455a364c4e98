#!/usr/bin/env python3
"""Helpers over `ncu --page source --csv` dumps (SASS view).
usage: ncu_src.py src.csv marks            -> landmark instructions with cumulative instruction/sample share
       ncu_src.py src.csv ops LO HI        -> opcode histogram of [LO,HI) (hex offsets), per warp given --warps N
       ncu_src.py src.csv dump LO HI       -> instructions of [LO,HI) with counts"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body, seen = [], set()
for r in rows[2:]:
    if len(r) > 10 and r[0].startswith("0x"):
        if r[0] in seen:
            break
        seen.add(r[0])
        body.append(r)
base = int(body[0][0], 16)
I = lambda r: int(r[ix["Instructions Executed"]])
S = lambda r: int(r[ix["# Samples"]])
tot, ts = sum(map(I, body)), sum(map(S, body))
cmd = sys.argv[2]
warps = 3125
if cmd == "marks":
    acc = accs = 0
    print("instructions %d samples %d" % (tot, ts))
    for r in body:
        off = int(r[0], 16) - base
        ins = r[1].strip()
        acc += I(r); accs += S(r)
        if any(t in ins for t in ["BAR.", "SYNCS", "UBLKCP", "RED", "ATOM", "VOTE", "SHFL", "WARPSYNC", "EXIT"]):
            print("%05x %-56s cum_inst=%5.1f%% cum_samp=%5.1f%% thr=%s" % (off, ins[:56], 100 * acc / tot, 100 * accs / ts, r[ix["Avg. Threads Executed"]]))
elif cmd == "ops":
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    h = collections.Counter(); t = 0; sm = 0
    for r in body:
        off = int(r[0], 16) - base
        if lo <= off < hi:
            ins = r[1].strip()
            op = (ins.split()[1] if ins.startswith("@") else ins.split()[0]).split(".")[0]
            h[op] += I(r); t += I(r); sm += S(r)
    print("region %x-%x: %d instr (%.1f%%), %.0f per warp, samples %.1f%%" % (lo, hi, t, 100 * t / tot, t / warps, 100 * sm / ts))
    for op, c in h.most_common(24):
        print("   %-10s %9d %5.1f%%  per-warp %7.0f" % (op, c, 100 * c / t, c / warps))
elif cmd == "dump":
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in body:
        off = int(r[0], 16) - base
        if lo <= off < hi:
            top = sorted(((int(r[ix[c]]), c[6:]) for c in st), reverse=True)[:2]
            print("%05x %-66s inst=%8d thr=%4s samp=%4d %s" % (off, r[1].strip()[:66], I(r), r[ix["Avg. Threads Executed"]], S(r), top))
