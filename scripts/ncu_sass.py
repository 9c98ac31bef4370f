#!/usr/bin/env python
"""From an `ncu --page source --csv --print-source cuda,sass` export: (1) a flat SASS listing with executed count and
stall samples per instruction (-> argv[2]), (2) samples by execution count (loop nest level), (3) the per-unit
(lo < count < hi) instructions grouped by the source line they are first listed under.
Usage: ncu_sass.py src.csv out_sass.txt [lo hi]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 0)


def f(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


hdr, cur, line, seen = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS, iA = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
        continue
    if hdr is None:
        continue
    if r[0] != "":
        line = (cur, f(r[0]), r[1].strip()[:80])
        continue
    if r[iA] in seen:
        continue
    seen[r[iA]] = (r[3].strip(), f(r[iI]), f(r[iS]), line)
open(sys.argv[2], 'w').write('\n'.join(f"{a[-5:]} {v[1]:8d} {v[2]:5d}  {v[0]}" for a, v in sorted(seen.items())) + '\n')
tot = sum(v[2] for v in seen.values())
cnt, smp = collections.Counter(), collections.Counter()
for v in seen.values():
    cnt[v[1]] += 1
    smp[v[1]] += v[2]
print('total samples', tot)
for k in sorted(cnt):
    if smp[k] > tot / 200:
        print(f"count {k:9d}: {cnt[k]:4d} instructions, {smp[k]:6d} samples ({100.0 * smp[k] / tot:4.1f} %)")
if hi:
    agg, ai = collections.Counter(), collections.Counter()
    for v in seen.values():
        if lo < v[1] < hi:
            agg[v[3]] += v[2]
            ai[v[3]] += 1
    for k, s in agg.most_common(32):
        print(f"{k[0]:20s} {k[1]:4d} i{ai[k]:3d} s{s:5d}  {k[2]}")
