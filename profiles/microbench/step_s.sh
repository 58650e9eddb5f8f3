#!/bin/bash
# builds one step_s binary per experiment value (here, no GPU needed) / runs them (on the GPU box)
cd "$(dirname "$0")"
VALUES="${VALUES:-0 1 2 3 4 8 16 32 63}"
if [ "$1" = build ]; then
    for v in $VALUES; do
        nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -std=c++17 -DSTAIR_EXP=$v -DIFL_STAIR_DEVICE_ONLY \
             -I ../../incremental-fluids_b200/csrc -I ../../include -o step_s_$v step_s.cu || exit 1
    done
else
    for v in $VALUES; do ./step_s_$v; done
fi
