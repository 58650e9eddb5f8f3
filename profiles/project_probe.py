"""Times FluidSolver::project(limit) (v3:349) alone on an N x N plume state, for engine / overlap experiments:
    IFL_TRI=2 IFL_OVERLAP_AXPY=3 python profiles/project_probe.py 16384 60 [height]
prints ms per PCG iteration (CUDA events around the whole solve, one warm solve first)."""
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ifl_binding", os.path.join(ROOT, "incremental-fluids_b200", "binding.py"))
binding = importlib.util.module_from_spec(spec)
spec.loader.exec_module(binding)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    h = int(sys.argv[3]) if len(sys.argv) > 3 else n
    torch.cuda.init()
    s = binding.FluidSolver(n, h, 0.1, version=3)
    s.addInflow(0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 3.0)
    s.buildRhs()
    s.buildPressureMatrix(0.005)
    s.buildPreconditioner()
    s.project(limit)
    s.sync()
    out = []
    for rep in range(2):
        s.buildRhs()
        s.sync()
        t0 = time.perf_counter()
        info = s.project(limit)
        s.sync()
        out.append((time.perf_counter() - t0) * 1e3 / limit)
    print("n=%dx%d limit=%d IFL_TRI=%s IFL_OVERLAP_AXPY=%s ms/iteration: %s  info=%s" % (
        n, h, limit, os.environ.get("IFL_TRI", "-"), os.environ.get("IFL_OVERLAP_AXPY", "-"), ["%.3f" % v for v in out], info))


if __name__ == "__main__":
    main()
