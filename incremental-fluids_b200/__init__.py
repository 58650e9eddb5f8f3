"""incremental-fluids_b200 -- B200 (sm_100a) implementation of the per-timestep hot
path of tunabrain/incremental-fluids, behind the reference's own FluidSolver surface.

Layout:
  csrc/              hand-written CUDA kernels + the C ABI (include/ifl_b200.h)
  libifl_b200.so     built in-tree by __graft_entry__.build() / csrc/Makefile
  host/              C++ drop-in classes (FluidSolver / FluidQuantity) over the C ABI
  binding.py         ctypes binding + a Python mirror of FluidSolver used by tests/bench

The directory name contains a hyphen (it is the reference's name), so import it with
    importlib.import_module("incremental-fluids_b200")
"""
from .binding import (  # noqa: F401
    BUF,
    FIELD,
    FluidSolver,
    IflError,
    KERNEL_CLASSES,
    SolidBox,
    SolidSphere,
    SolveInfo,
    build_library,
    library_path,
    load_library,
)
