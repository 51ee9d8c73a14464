#!/bin/bash
# A/B the library variants in gpurun_variants/ on the index build (sort / rle / table ms of the third build).
for lib in gpurun_variants/lib_*.so; do
  echo -n "$lib  "; PA_B200_LIB=$PWD/$lib python tools/build_only.py 2>&1 | tail -1
done
