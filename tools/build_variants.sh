#!/bin/bash
# Builds library variants that differ in the compile-time knobs of one translation unit (default align.cu; SRC=sort.cu ...)
# into gpurun_variants/ (A/B them on the box with tools/ab.sh / tools/ab_build.sh).
# usage: [SRC=sort.cu] tools/build_variants.sh name1:"-DFLAG=.. -DFLAG2=.." name2:"..." ...
set -e
SRC=${SRC:-align.cu}
P=bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200
mkdir -p gpurun_variants build/obj
python -c "import __graft_entry__ as g; g._compile(False)"
OTHERS=$(ls build/obj/*.o | grep -v "/$SRC.o")
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-pthread $flags -c $P/csrc/$SRC -o build/obj/variant_$name.o &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC,-pthread -o gpurun_variants/lib_$name.so build/obj/variant_$name.o $OTHERS -ldl
  rm -f build/obj/variant_$name.o
done
ls -la gpurun_variants/
