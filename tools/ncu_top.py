#!/usr/bin/env python
"""Summarise an ncu report: key metrics + the most-sampled SASS instructions (needs -lineinfo / --import-source).
usage: python tools/ncu_top.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
if len(rows) >= 3:
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")][:80])
        for key in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
                    "launch__registers_per_thread", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"):
            if key in hdr:
                i = hdr.index(key)
                print(f"   {key:75s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
# the first line is the kernel name row; the header is the second
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(lines[start:])
recs = []
for r in rd:
    try:
        recs.append((int(r["# Samples"]), r["Source"].strip(), r))
    except (KeyError, ValueError):
        continue
total = sum(s for s, _, _ in recs) or 1
print(f"-- top {topn} sampled instructions of {len(recs)} (total samples {total})")
stall_cols = [c for c in rd.fieldnames if c.startswith("stall_") and "Not Issued" not in c]
for idx, (s, text, r) in sorted(enumerate(recs), key=lambda t: -t[1][0])[:topn]:
    st = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100.0 * s / total:5.1f}%  #{idx:5d}  {text[:70]:70s} {st}")
