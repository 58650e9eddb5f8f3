"""Workload for the ncu captures (profiles/ncu_capture.sh): one short run of the chapter whose kernels are
being profiled.  Numbers printed under ncu are never bench values."""
import importlib
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ifl = importlib.import_module("incremental-fluids_b200")
chapter = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
if chapter == 8:
    s = ifl.FluidSolver(n, n, 0.1, version=8, bodies=[ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0)],
                        rho_soot=0.25, diffusion=0.01, avg_per_cell=8)
    s.update(0.0025)
elif chapter == 2:
    s = ifl.FluidSolver(n, n, 0.1, version=2)
    s.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    s.buildRhs()
    s.project(24, 0.005)
    for k in "duv":
        s.advect(k, 0.005)
else:
    s = ifl.FluidSolver(n, n, 0.1, version=3)
    s.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    s.buildRhs()
    s.buildPressureMatrix(0.005)
    s.buildPreconditioner()
    s.project(12)
    s.applyPressure(0.005)
    for k in "duv":
        s.advect(k, 0.005)
s.sync()
print("ncu target done: chapter %d, %dx%d" % (chapter, n, n))
