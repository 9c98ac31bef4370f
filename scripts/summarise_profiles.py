#!/usr/bin/env python
"""Turns the files scripts/capture_profiles.sh left in gpurun_out/ into what is tracked under profiles/:
raw CSVs and bench lines copied, launch list trimmed + per-kernel shares, opcode histograms, SASS mnemonic evidence,
and a few per-kernel key figures printed for profiles/README.md.   python scripts/summarise_profiles.py"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNELS = (("fused_tile", "ekf_fused_tile_kernel<false>"), ("sweep_mma", "k_large_sweep_mma<14>"), ("circles", "k_circles_scan<float>"))

for f in ("r2_bench_default_n1.json", "r2_bench_reference_arm.json") + tuple(f"r2_prof_{n}_raw.csv" for n, _ in KERNELS):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))

# ---- launch list
rows = list(csv.reader(l for l in open(os.path.join(G, "r2_launches_bench.csv")) if l.startswith('"')))
hdr = rows[0]
iN, iV, iM, iU, iG, iB, iI = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "Grid Size", "Block Size", "ID"))
agg, total, out = collections.OrderedDict(), 0.0, [("ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum [us]")]
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    us = {"nsecond": v / 1e3, "ns": v / 1e3, "usecond": v, "us": v, "msecond": v * 1e3, "ms": v * 1e3}.get(r[iU], v)
    a = agg.setdefault(r[iN], [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    out.append((r[iI], r[iN], r[iG], r[iB], f"{us:.3f}"))
with open(os.path.join(P, "r2_launches_bench.csv"), "w", newline="") as f:
    csv.writer(f).writerows(out)
with open(os.path.join(P, "r2_launches_bench_summary.txt"), "w") as f:
    f.write(f"{len(out) - 1} launches, {total / 1e3:.1f} ms of device time (ncu, serialised, cold)\n")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t / total * 100:6.2f} %  {n:6d} x {t / n:10.2f} us  {name[:110]}\n")
print(open(os.path.join(P, "r2_launches_bench_summary.txt")).read())

# ---- opcode histograms + SASS evidence
ev = []
for n, kern in KERNELS:
    src = os.path.join(G, f"r2_prof_{n}_source.csv")
    with open(os.path.join(P, f"r2_opcodes_{n}.txt"), "w") as f:
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_opcodes.py"), src, "25"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(open(src)))
    hdr, seen, static, dyn = None, set(), collections.Counter(), collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "Line No":
            hdr = r
            iE, iA = hdr.index("Instructions Executed"), hdr.index("Address")
            continue
        if hdr is None or r[0] != "" or len(r) <= iE or r[iA] in seen:
            continue
        seen.add(r[iA])
        parts = r[3].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        try:
            ex = int(float(r[iE]))
        except ValueError:
            ex = 0
        static[op] += 1
        dyn[op] += ex
    ev.append(f"== {kern}: {len(seen)} SASS instructions in the kernel image (static), executed warp instructions in brackets")
    for key in ("DMMA", "UBLKCP", "SYNCS", "REDUX", "LDS", "STS", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG"):
        items = [(op, c, dyn[op]) for op, c in static.items() if op.split(".")[0] == key]
        ev.append(f"  {key:8s} " + ("; ".join(f"{op} x{c} [{d}]" for op, c, d in sorted(items, key=lambda x: -x[1])[:6]) if items else "none"))
    ev.append("")
open(os.path.join(P, "r2_sass_evidence.txt"), "w").write("\n".join(ev))

# ---- key figures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active"]
for n, kern in KERNELS:
    rows = list(csv.reader(open(os.path.join(P, f"r2_prof_{n}_raw.csv"))))
    hdr, units = rows[0], rows[1]
    print("==", kern)
    for r in rows[2:3]:
        for w in WANT:
            if w in hdr:
                print(f"   {w}: {r[hdr.index(w)]} {units[hdr.index(w)]}")
