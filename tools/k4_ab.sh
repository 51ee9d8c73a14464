#!/bin/bash
# control experiment: round-1 tree (build/old) against the current tree, same box, same workload; K4 time + ncu DRAM/instruction counters
mkdir -p gpurun_out
for T in old new; do
  if [ $T = old ]; then D=build/old; else D=.; fi
  ( cd $D && python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'tree': '$T', 'k4_ms': d['roofline']['kernel_ms'], 'ms_per_step': d['ms_per_step'], 'build': d['build'], 'clocks': d['clocks']}))
" ) >> gpurun_out/k4_ab.jsonl
  ( cd $D && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_op_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:align_fast -c 2 --csv --log-file $OLDPWD/gpurun_out/k4_ab_ncu_$T.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 )
done
cat gpurun_out/k4_ab.jsonl
