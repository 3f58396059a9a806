#!/usr/bin/env python
"""Top CUDA source lines by executed warp-instructions / stall samples from an .ncu-rep (needs -lineinfo).
Usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = func = None
hdr = None
acc = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1].split("(")[0]
        continue
    if r[0] == "Line No":
        hdr = r
        iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit():
        try:
            ex, sa = int(r[iex]), int(r[isamp])
        except ValueError:
            continue
        key = (func, fname, int(r[0]), r[1].strip())
        a = acc.setdefault(key, [0, 0])
        a[0] += ex
        a[1] += sa
by_func = {}
for (func, fname, ln, src), (ex, sa) in acc.items():
    by_func.setdefault(func, []).append((ex, sa, fname, ln, src))
for func, lst in by_func.items():
    tot = sum(x[0] for x in lst)
    ts = sum(x[1] for x in lst)
    print(f"=== {func}: executed warp-instructions (all profiled launches) {tot}, stall samples {ts}")
    for ex, sa, fname, ln, src in sorted(lst, reverse=True)[:top]:
        print(f"  {100*ex/max(tot,1):5.1f}% inst  {100*sa/max(ts,1):5.1f}% stall  {fname}:{ln}: {src[:95]}")
