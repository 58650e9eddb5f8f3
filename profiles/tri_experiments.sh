#!/bin/bash
# Builds one libifl_b200 per TRI_EXP value (timing experiments of tri_kernels.cu) into profiles/exp/
# and, with "run", prints strip 0's cycles per step for each of them (B200 box).
set -e
cd "$(dirname "$0")/.."
SRC=incremental-fluids_b200/csrc
OUT=profiles/exp
mkdir -p $OUT
VARIANTS="${VARIANTS:-0 1 2 4 8 6 14}"
if [ "$1" != "run" ]; then
  make -C $SRC -j8 >/dev/null
  for v in $VARIANTS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -DTRI_EXP=$v $EXTRA -c -o $OUT/tri_$v.o $SRC/tri_kernels.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libifl_exp$v.so $SRC/ifl_api.o $SRC/dist.o $SRC/pcg_kernels.o $SRC/sweep_kernels.o $OUT/tri_$v.o $SRC/assembly_kernels.o $SRC/advect_kernels.o $SRC/solid_kernels.o $SRC/flip_kernels.o
  done
else
  for v in $VARIANTS; do
    echo "=== TRI_EXP=$v"
    IFL_B200_LIB=$PWD/$OUT/libifl_exp$v.so timeout 120 python profiles/sweep_timeline.py ${SIZE:-4096} 2>&1 | sed -n 2,7p
  done
fi
