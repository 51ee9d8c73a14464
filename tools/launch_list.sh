#!/bin/bash
# Launch list (ncu --metrics gpu__time_duration.sum, per B200_PROFILING.md) of a short bench run -- three single-GPU builds
# and the align steps -- own kernels only, after the same command has exited 0 without ncu.
# usage: tools/launch_list.sh out.csv [bench args]
OUT=${1:-gpurun_out/launches.csv}; shift
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs $*"
KERNELS='regex:align_fast_|align_kernel|quality_masks|summary_kernel|radix_|fix_|rle_|encode_windows|table_insert|stash_insert|long_c|set_heads|set_assign|scan_u64|tile_sum|tile_scan|set_csr|scratch_init|first_occ|iota|owner_|genome_map|slice_|bitmap_|csr_checksum'
$CMD > gpurun_out/launch_list_plain.log 2>&1 || { tail -5 gpurun_out/launch_list_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv --log-file $OUT $CMD > gpurun_out/launch_list_ncu.log 2>&1
grep -v "^==" $OUT | awk -F'","' '{print substr($5,1,64), $NF}' | tr -d '"' | awk '{n=$NF; $NF=""; t[$0]+=n; c[$0]++} END {for (k in t) printf "%-66s x%-3d %10.3f ms\n", k, c[k], t[k]/1e6}' | sort -k1
