#!/bin/bash
# A/B the library variants in gpurun_variants/ on the bench kernel time.  Usage: tools/ab.sh [bench args]
for lib in gpurun_variants/lib_*.so; do
  PA_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'kernel_ms %.2f' % d['roofline']['kernel_ms'], 'step_ms %.2f' % d['ms_per_step'], d['result'])"
done
