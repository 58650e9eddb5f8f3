"""Device timings of the other BASELINE.json configurations (1 GPU), at sizes a short run allows:
  config 2: chapter 2 (Gauss-Seidel, limit 600) and chapter 3 (PCG, limit 600) at 2048^2, 3 updates each
  config 3: chapter 5 at 2048^2 with the rotating box / sphere / box of SURVEY 8d, limit 2000, 2 updates
  config 4: chapter 7 (heat + variable density) at 2048^2, 2 updates
Prints one JSON object; `python profiles/config_runs.py > profiles/rNN_configs.json` under gpurun."""
import importlib
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ifl = importlib.import_module("incremental-fluids_b200")


def run(name, solver, steps, inflow, move=None):
    solver.addInflow(*inflow)
    solver.update(0.005)  # warm-up
    solver.sync()
    solver.profile(True)
    t0 = time.perf_counter()
    infos = []
    for i in range(steps):
        solver.addInflow(*inflow)
        infos.append(solver.update(0.005))
        if move and (i % 4) == 3:
            move()
    solver.sync()
    dt = time.perf_counter() - t0
    prof = {k: v for k, v in solver.profile_read().items() if v[1] > 0}
    solver.profile(False)
    cells = solver.w * solver.h
    out = {"config": name, "grid": [solver.w, solver.h], "steps": steps, "s_per_step": dt / steps,
           "cell_updates_per_s": cells * steps / dt, "solver": [list(x) for x in infos],
           "kernel_ms": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in prof.items()}}
    if "gs_sweep" in prof:
        out["gs_sweeps_per_s"] = prof["gs_sweep"][1] / (prof["gs_sweep"][0] * 1e-3)
    if "matvec" in prof:
        pcg = sum(prof[k][0] for k in ("matvec", "axpy2_norm", "precon_fwd", "precon_bwd", "xpay", "scalar") if k in prof)
        out["pcg_iters_per_s"] = prof["matvec"][1] / (pcg * 1e-3)
    solver.close()
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    res = []
    res.append(run("2: chapter 2, Gauss-Seidel", ifl.FluidSolver(n, n, 0.1, version=2), 3, (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)))
    res.append(run("2: chapter 3, MIC(0)-PCG", ifl.FluidSolver(n, n, 0.1, version=3), 3, (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)))
    bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 1.0),
              ifl.SolidSphere(0.15, 0.3, 0.15, 0.0, 0.0, 0.0, 0.0),
              ifl.SolidBox(0.85, 0.2, 0.2, 0.1, 0.0, 0.0, 0.0, -2.0)]
    s5 = ifl.FluidSolver(n, n, 0.1, version=5, bodies=bodies)
    res.append(run("3: chapter 5, rotating bodies", s5, 2, (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0),
                   move=lambda: [b.update(0.005) for b in bodies]))
    b7 = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0)]
    s7 = ifl.FluidSolver(n, n, 0.1, version=7, bodies=b7, rho_soot=1.0, diffusion=0.01)
    res.append(run("4: chapter 7, heat + variable density", s7, 2, (0.45, 0.2, 0.15, 0.03, 1.0, s7.ambientT() + 300.0, 0.0, 0.0)))
    print(json.dumps({"configs": res}, indent=1))


if __name__ == "__main__":
    main()
