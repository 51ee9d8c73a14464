#!/bin/bash
# K4 timings of the current build: plain and EXTQUALITY (mask pre-kernel + staggered K4, and the in-kernel quality scan); one JSON line each into gpurun_out/k4_variants.jsonl
mkdir -p gpurun_out
: > gpurun_out/k4_variants.jsonl
run() {
  label=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs $EXTRA 2>gpurun_out/k4_variants_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'variant': '$label', 'k4_ms': d['roofline']['kernel_ms'], 'ms_per_step': d['ms_per_step'], 'frac': d['roofline']['frac'], 'digest_ok': (d.get('parity_fullsize_digest') or {}).get('equals_oracle_digest'), 'result': d['result']}))
" >> gpurun_out/k4_variants.jsonl
}
EXTRA="" run plain PA_X=0
EXTRA="--extquality" run extq_masks PA_QUAL_MASKS=1
EXTRA="--extquality" run extq_in_kernel PA_QUAL_MASKS=0
cat gpurun_out/k4_variants.jsonl
