#!/bin/bash
# K4 time and table size against the table's load factor (config B); one JSON line per load into gpurun_out/load_sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/load_sweep.jsonl
for L in ${LOADS:-0.20 0.25 0.30 0.36 0.42}; do
  PA_TABLE_LOAD=$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>gpurun_out/load_sweep_err_$L.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'load': $L, 'k4_ms': d['roofline']['kernel_ms'], 'ms_per_step': d['ms_per_step'], 'table_bytes': d['build']['table_bytes'], 'stash': d['build']['stash_count'], 'table_ms': d['build']['table_ms'], 'bytes_per_kmer': d['build']['table_bytes'] / d['build']['distinct_kmers'], 'result': d['result']}))
" >> gpurun_out/load_sweep.jsonl
done
cat gpurun_out/load_sweep.jsonl
