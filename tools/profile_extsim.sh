#!/bin/bash
CMD="python tools/bench_extsim.py --genomes 300 --genome-len 1000000"
KERNELS='regex:radix_|rle_|encode_windows|table_insert|stash_insert|mlist_fill|msector_counts|scan_u64|tile_sum|tile_scan|set_csr|first_occ|iota|extsim|drop_'
$CMD > gpurun_out/extsim_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 300 --csv --log-file gpurun_out/launches_extsim.csv $CMD > gpurun_out/ncu_extsim.log 2>&1
tail -1 gpurun_out/extsim_plain.log | cut -c1-900
