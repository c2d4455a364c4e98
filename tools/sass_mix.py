#!/usr/bin/env python3
"""Static opcode mix of one kernel in libhulk_b200.so (cuobjdump -sass): all instructions, and the instructions of
every backward-branch loop body (label .. the branch back to it), innermost first.  No GPU needed; the dynamic counts are
in profiles/*_ncu_full_summary.txt.  usage: sass_mix.py <substring of the mangled kernel name> [lib]"""
import collections, os, re, subprocess, sys

name = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hulk_b200", "libhulk_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next((f for f in funcs[1:] if name in f.split("\n", 1)[0]), None)
if body is None:
    sys.exit("no kernel matching %r" % name)
print("kernel:", body.split("\n", 1)[0])
ins = []                                          # (addr, opcode, text)
for ln in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m:
        continue
    t = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
    full = t.split()[0]
    op = full.split(".")[0]
    if op == "IMAD":                                  # IMAD.MOV / IMAD.SHL / IMAD.IADD are moves, shifts and adds in disguise
        sub = full.split(".")[1] if "." in full else ""
        op = "IMAD." + sub if sub in ("MOV", "SHL", "IADD", "WIDE", "HI", "X") else "IMAD"
    ins.append((int(m.group(1), 16), op, t))


def mix(rows):
    c = collections.Counter(op for _, op, _ in rows)
    return ", ".join("%s %d" % kv for kv in c.most_common(14))


print("all: %d instructions: %s" % (len(ins), mix(ins)))
loops = []
for addr, op, t in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
for lo, hi in sorted(loops, key=lambda p: p[1] - p[0]):
    rows = [r for r in ins if lo <= r[0] <= hi]
    print("loop 0x%04x..0x%04x: %d instructions: %s" % (lo, hi, len(rows), mix(rows)))
