#!/bin/bash
# ncu --set full of the two heaviest build kernels (K2 radix_pass, table_insert) on the bench workload's build.
set -u
CMD="python tools/build_only.py"
$CMD > gpurun_out/build_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:radix_pass -s 1 -c 1 -f -o gpurun_out/prof_radix $CMD > gpurun_out/ncu_radix.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:table_insert -s 1 -c 1 -f -o gpurun_out/prof_insert $CMD > gpurun_out/ncu_insert.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rle_scatter -s 1 -c 1 -f -o gpurun_out/prof_rle $CMD > gpurun_out/ncu_rle.log 2>&1
tail -2 gpurun_out/build_plain.log; tail -2 gpurun_out/ncu_radix.log gpurun_out/ncu_insert.log gpurun_out/ncu_rle.log
