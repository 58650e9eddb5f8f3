#!/bin/bash
# One `ncu --set full --clock-control none` capture per kernel class of the hot path (B200 box, ONE GPU):
#   chapter 3 at 4096^2: k_tri (forward, backward + dot), k_matvec, k_xpay_matvec, k_axpy2_norm, k_scaled_add, k_advect,
#                        k_sweep (MIC(0) factorisation), k_build_*, k_apply_pressure_*
#   chapter 2 at 2048^2: k_sweep (Gauss-Seidel)
#   chapter 8 at 1024^2: k_from_particles, k_grid_to_particles, k_particles_advect
# Reports land in gpurun_out/ncu_r02/; profiles/ncu_summarise.py turns them into the committed table.
cd "$(dirname "$0")/.."
OUT=gpurun_out/ncu_r02
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { # name regex skip count chapter size
  timeout 600 $NCU -k "regex:$2" --launch-skip $3 --launch-count $4 -o $OUT/$1 python profiles/ncu_target.py $5 $6 > $OUT/$1.log 2>&1 || echo "capture $1 failed"
}
cap tri           '^k_tri'            6 2 3 4096
cap matvec        '^k_matvec'         0 1 3 4096
cap xpay_matvec   '^k_xpay_matvec'    4 1 3 4096
cap axpy2_norm    '^k_axpy2_norm'     4 1 3 4096
cap scaled_add    '^k_scaled_add'     0 1 3 4096
cap advect        '^k_advect'         0 1 3 4096
cap factor        '^k_sweep'          0 1 3 4096
cap assembly      '^k_build|^k_apply' 0 4 3 4096
cap gs_sweep      '^k_sweep'          4 1 2 2048
cap p2g           '^k_from_particles' 0 1 8 1024
cap g2p           '^k_grid_to_particles' 4 1 8 1024
cap padvect       '^k_particles_advect'  0 1 8 1024
ls -la $OUT
