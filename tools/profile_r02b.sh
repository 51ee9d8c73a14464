#!/bin/bash
# Final round-2 captures on one GPU (each after its command exited 0 without ncu): launch lists of the plain and EXTQUALITY
# runs, ncu --set full of K4 (plain), of K4 on quality masks + quality_masks_kernel (EXTQUALITY), and of the build's
# heaviest kernels (radix_pass, table_insert, rle_scatter, fix_detect, long_collect).
set -u
bash tools/launch_list.sh gpurun_out/r02_launches_v3.csv > gpurun_out/r02_launches_v3_summary.txt
bash tools/launch_list.sh gpurun_out/r02_launches_extq_v3.csv --extquality > gpurun_out/r02_launches_extq_v3_summary.txt
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
ncu --set full --clock-control none --import-source on -k regex:align_fast_ -s 3 -c 1 -f -o gpurun_out/r02_align_fast_v3 $CMD > gpurun_out/ncu_k4.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_align_fast_v3.ncu-rep 70 > gpurun_out/r02_align_fast_ncu_v3_lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/r02_align_fast_v3.ncu-rep > gpurun_out/r02_align_fast_ncu_v3.json 2>gpurun_out/ncu_summary.err
ncu --set full --clock-control none --import-source on -k regex:'align_fast_|quality_masks' -s 6 -c 2 -f -o gpurun_out/r02_extq_v3 $CMD --extquality > gpurun_out/ncu_extq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'radix_pass|table_insert|rle_scatter|fix_detect|long_collect' -s 4 -c 9 -f -o gpurun_out/r02_build_v3 $CMD > gpurun_out/ncu_build.log 2>&1
: > gpurun_out/r02_kernels_ncu_v3.jsonl
for rep in r02_extq_v3 r02_build_v3; do
  ncu -i gpurun_out/$rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys, json
rows = list(csv.reader(sys.stdin))
if len(rows) < 3: sys.exit(0)
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
seen = set()
for r in rows[2:]:
    name = r[h.index('Kernel Name')].split('(')[0]
    if name in seen: continue
    seen.add(name)
    print(json.dumps({'report': '$rep', **{k: r[h.index(k)] for k in want if k in h}}))
" >> gpurun_out/r02_kernels_ncu_v3.jsonl
done
cat gpurun_out/r02_launches_v3_summary.txt; grep -i "align\|quality" gpurun_out/r02_launches_extq_v3_summary.txt; cat gpurun_out/r02_kernels_ncu_v3.jsonl | cut -c1-400; head -8 gpurun_out/r02_align_fast_ncu_v3_lines.txt | cut -c1-200
