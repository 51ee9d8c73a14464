#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture of the top kernel.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs $*"
KERNELS='regex:align_fast_|align_kernel|summary_kernel|radix_|rle_|encode_windows|table_insert|stash_insert|long_flags|set_hash_keys|set_heads|set_assign|scan_u64|tile_sum|tile_scan|set_csr|scratch_init|first_occ|iota|owner_'
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:align_fast_ -s 3 -c 1 -f -o gpurun_out/prof_align $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log | cut -c1-600
tail -3 gpurun_out/ncu_full.log
