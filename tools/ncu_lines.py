#!/usr/bin/env python3
"""Per-CUDA-source-line stall samples of an .ncu-rep.  Usage: python tools/ncu_lines.py report.ncu-rep [top=40]"""
import csv, io, subprocess, sys

def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname = ""
lines = []
h = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if "# Samples" in r:
        h = r
        ci = {}
        for i, n in enumerate(h):
            ci.setdefault(n, i)
        continue
    if h is None or len(r) != len(h) or not r[0].strip().isdigit():
        continue
    lines.append((fname, r))
tot = sum(num(r[ci["# Samples"]]) for _, r in lines)
print("total samples", tot)
for f, r in sorted(lines, key=lambda x: -num(x[1][ci["# Samples"]]))[:top_n]:
    n = num(r[ci["# Samples"]])
    if n == 0:
        break
    def pc(c):
        return num(r[ci[c]]) * 100 // max(n, 1)
    print(f"{f}:{r[0]:>4} {100 * n / tot:5.2f}% inst={num(r[ci['Instructions Executed']]) / 1e6:8.1f}M "
          f"long_sb={pc('stall_long_sb'):3d}% no_inst={pc('stall_no_inst'):3d}% wait={pc('stall_wait'):3d}% "
          f"short_sb={pc('stall_short_sb'):3d}% branch={pc('stall_branch_resolving'):3d}% | {r[1].strip()[:90]}")
