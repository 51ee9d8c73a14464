#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel) into key metrics + stall breakdown + hottest source lines.
Usage: python tools/ncu_summary.py gpurun_out/prof_align.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "Kernel Name"]
out = {"report": rep, "metrics": {}}
for i, h in enumerate(hdr):
    if h in want:
        out["metrics"][h] = {"value": vals[i], "unit": units[i]}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(num(r[ci["# Samples"]]) for r in data)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(num(r[ci[s]]) for r in data) for s in stalls}
out["samples"] = tot
out["stalls_pct"] = {s: round(100 * v / tot, 2) for s, v in sorted(agg.items(), key=lambda x: -x[1]) if v}
srccol = "Source" if "Source" in ci else h[1]
top = sorted(data, key=lambda r: -num(r[ci["# Samples"]]))[:40]
out["hot"] = [{"src": r[ci[srccol]].strip()[:110], "samples_pct": round(100 * num(r[ci["# Samples"]]) / tot, 2),
               "inst_executed": r[ci["Instructions Executed"]]} for r in top]
js = json.dumps(out, indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(js)
print(js)
