#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) of one kernel in an .ncu-rep.
usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hd = rows[h]
ie, ws = hd.index("Instructions Executed"), hd.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows[h + 1:]:
    if r and r[0].strip().isdigit():
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[ws] or 0)))
        except ValueError:
            pass
ti, ts = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
print(f"total warp-instr {ti}  stall samples {ts}")
for ln, src, n, s in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{ln:5d} {100*n/ti:5.1f}% instr {100*s/ts:5.1f}% stall  {src[:110]}")
