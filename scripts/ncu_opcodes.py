#!/usr/bin/env python
"""Opcode histogram (executed warp instructions, stall samples, shared-memory wavefronts) of an
`ncu --page source --csv --print-source cuda,sass` export.  Usage: ncu_opcodes.py src.csv [top] [per_unit_divisor]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
div = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
inst = collections.Counter()
samp = collections.Counter()
wave = collections.Counter()
hdr = None
seen = set()
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS, iW, iA = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared"), hdr.index("Address")
        continue
    if hdr is None or r[0] != "" or len(r) <= iW:
        continue
    addr = r[iA]
    if addr in seen:  # a SASS instruction is listed under every source line it is attributed to
        continue
    seen.add(addr)
    sass = r[3].strip()
    parts = sass.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "LDL", "STL")) else op.split(".")[0]
    try:
        n, s, w = int(float(r[iI] or 0)), int(float(r[iS] or 0)), int(float(r[iW] or 0))
    except ValueError:
        continue
    inst[op] += n
    samp[op] += s
    wave[op] += w
tot, ts, tw = sum(inst.values()), sum(samp.values()), sum(wave.values())
print(f"total warp instructions {tot}  samples {ts}  shared wavefronts {tw}" + (f"  per unit: {tot / div:.1f} inst, {tw / div:.1f} wavefronts" if div else ""))
for op, n in inst.most_common(top):
    extra = f"  {n / div:7.1f}/unit" if div else ""
    print(f"{op:14s} {n / tot * 100:5.1f}%  samples {samp[op] / max(ts, 1) * 100:5.1f}%  wavefronts {wave[op] / max(tw, 1) * 100:5.1f}%{extra}")
