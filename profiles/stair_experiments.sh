#!/bin/bash
# Builds one libifl_b200 per STAIR_EXP value (timing experiments of stair_kernels.cu) into profiles/exp/
# and, with "run", prints strip 0's cycles per step and the sweep time for each of them (B200 box).
set -e
cd "$(dirname "$0")/.."
SRC=incremental-fluids_b200/csrc
OUT=profiles/exp
mkdir -p $OUT
VARIANTS="${VARIANTS:-0 64 128 192}"
if [ "$1" != "run" ]; then
  make -C $SRC -j8 >/dev/null
  for v in $VARIANTS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -DSTAIR_EXP=$v $EXTRA -c -o $OUT/stair_$v.o $SRC/stair_kernels.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libifl_sexp$v.so $(ls $SRC/*.o | grep -v stair_kernels.o) $OUT/stair_$v.o
  done
else
  for v in $VARIANTS; do
    echo "=== STAIR_EXP=$v"
    IFL_TRI=2 IFL_B200_LIB=$PWD/$OUT/libifl_sexp$v.so timeout 120 python profiles/sweep_timeline.py ${SIZE:-4096} 2>&1 | sed -n 2,7p
  done
fi
