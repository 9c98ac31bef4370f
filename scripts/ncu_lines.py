#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line (instructions, samples)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.Counter(); samp = collections.Counter(); srcs = {}
cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and cur:
        try:
            ln = int(r[0]); n = int(float(r[iI] or 0)); s = int(float(r[iS] or 0))
        except Exception:
            continue
        if n or s:
            agg[(cur, ln)] += n; samp[(cur, ln)] += s; srcs[(cur, ln)] = r[1].strip()[:100]
tot = sum(agg.values()); ts = sum(samp.values())
print("total inst", tot, "samples", ts)
for (f, l), n in agg.most_common(top):
    print(f"{f:22s} {l:4d} {n/tot*100:5.1f}% samp {samp[(f,l)]/max(ts,1)*100:5.1f}%  {srcs[(f,l)]}")
