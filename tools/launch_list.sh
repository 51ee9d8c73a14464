#!/bin/bash
# Launch list (ncu --metrics gpu__time_duration.sum, per B200_PROFILING.md) of one single-GPU build + the align steps of a
# short bench run, own kernels only.  usage: tools/launch_list.sh out.csv [bench args]
OUT=${1:-gpurun_out/launches.csv}; shift
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs $*"
$CMD > gpurun_out/launch_list_plain.log 2>&1 || { tail -5 gpurun_out/launch_list_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!.*(at::|at_cuda|cub::|elementwise|distribution)).*$' -c 600 --csv --log-file $OUT $CMD > gpurun_out/launch_list_ncu.log 2>&1
grep -v "^==" $OUT | awk -F'","' '{print substr($5,1,64), $NF}' | tr -d '"' | awk '{n=$NF; $NF=""; t[$0]+=n; c[$0]++} END {for (k in t) printf "%-66s x%-3d %10.3f ms\n", k, c[k], t[k]/1e6}' | sort -k4 -n -r | head -50
