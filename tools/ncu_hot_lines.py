#!/usr/bin/env python
"""Per-source-line hot spots from an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_hot_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        if hdr is not None:
            break  # first launch only
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        lines.append(r)
ix = {h: i for i, h in enumerate(hdr)}
I, S, T = ix["Instructions Executed"], ix["# Samples"], ix["Thread Instructions Executed"]
f = lambda v: float(v) if v not in ("", "-") else 0.0
tot_i = sum(f(r[I]) for r in lines) or 1
tot_s = sum(f(r[S]) for r in lines) or 1
print(f"kernel {kern}: {tot_i:.3g} warp instructions, {tot_s:.0f} samples (first matching launch)")
for r in sorted(lines, key=lambda r: -f(r[S]))[:top]:
    thr = f(r[T]) / f(r[I]) if f(r[I]) else 0
    print(f"L{r[0]:>4s} {f(r[S]) / tot_s * 100:5.1f}% smp {f(r[I]) / tot_i * 100:5.1f}% inst {thr:4.1f} thr | {r[1].strip()[:105]}")

if len(sys.argv) > 4 and sys.argv[4] == "ranges":
    # cumulative instruction share per 10-line bucket, in line order
    print("\nline-ordered instruction share (lines with >= 0.3 %):")
    for r in sorted(lines, key=lambda r: int(r[0])):
        if f(r[I]) / tot_i * 100 >= 0.3:
            print(f"L{r[0]:>4s} {f(r[I]) / tot_i * 100:5.1f}% inst {f(r[S]) / tot_s * 100:5.1f}% smp | {r[1].strip()[:100]}")
