"""One rank of a row-slab multi-GPU run (launched by test_gpu_dist.py / by hand):
    RANK=r WORLD_SIZE=n python tests/dist_worker.py <rendezvous> <chapter> <w> <h> <steps>
Runs the plume on `world` GPUs through the C ABI, then the same plume on ONE GPU (this
rank's), and demands bit-identical fields, solver status and iteration counts; rank 0 also
checks against the CPU oracle (<= 1e-10, equal iteration counts) at sizes it finishes fast."""
import importlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ifl = importlib.import_module("incremental-fluids_b200")


def make_solver(chapter, w, h, **kw):
    """The chapter's shipped plume: chapters 4+ with a rotating box and a sphere (the solver
    holds the list by reference, the caller moves the bodies; v5:986-1011), chapters 6-7 with
    the heat / variable-density constructor (v7:1091-1099)."""
    bodies = None
    if chapter >= 4:
        bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.3),
                  ifl.SolidSphere(0.3, 0.25, 0.12, 0.0, 0.0, 0.0, 0.0)]
    if chapter >= 6:
        return ifl.FluidSolver(w, h, 0.1, version=chapter, bodies=bodies, rho_soot=1.0 if chapter == 7 else 0.1,
                               diffusion=0.01, **kw)
    return ifl.FluidSolver(w, h, 0.1, version=chapter, bodies=bodies, **kw)


def plume(solver, steps, chapter):
    infos = []
    for i in range(steps):
        if chapter >= 6:
            solver.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, solver.ambientT() + 300.0, 0.0, 0.0)  # v6:1086
        else:
            solver.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        infos.append(solver.update(0.005))
        if chapter >= 6:
            infos.append(solver.last_heat)
        if chapter >= 4:
            for b in solver.bodies:
                b.update(0.005)
    return infos


def main():
    rdv, chapter, w, h, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    multi = make_solver(chapter, w, h, device=rank, rank=rank, world=world, rendezvous=rdv)
    r0, r1 = multi.rows()
    assert r0 % 64 == 0 and r0 < r1 <= h, (r0, r1)
    rng = np.random.default_rng(7)
    rvec = rng.uniform(-1, 1, w * h)
    gran = {}
    if chapter == 3:  # granular hot-path ops across the slab boundary
        multi.buildPressureMatrix(0.005)
        multi.buildPreconditioner()
        multi.set("r", rvec)
        multi.applyPreconditioner("z", "r")
        multi.matrixVectorProduct("s", "z")
        gran = {"z": multi.get("z"), "s": multi.get("s"), "precon": multi.get("precon"),
                "dot": multi.dotProduct("z", "r"), "norm": multi.infinityNorm("s")}
        multi.set("r", np.zeros(w * h))
    infos_m = plume(multi, steps, chapter)
    names = ["d.src", "u.src", "v.src", "p"] + (["t.src"] if chapter >= 6 else [])
    got = {k: multi.get(k) for k in names}
    launches = multi.launches()
    multi.barrier()

    single = make_solver(chapter, w, h, device=rank)
    if chapter == 3:
        single.buildPressureMatrix(0.005)
        single.buildPreconditioner()
        single.set("r", rvec)
        single.applyPreconditioner("z", "r")
        single.matrixVectorProduct("s", "z")
        for k in ("z", "s", "precon"):
            assert np.array_equal(gran[k], single.get(k)), "rank %d: %s differs from the one-GPU run" % (rank, k)
        assert gran["dot"] == single.dotProduct("z", "r"), "dot differs"
        assert gran["norm"] == single.infinityNorm("s"), "norm differs"
        single.set("r", np.zeros(w * h))
    infos_s = plume(single, steps, chapter)
    assert infos_m == infos_s, (infos_m, infos_s)
    for k in names:
        a, b = got[k], single.get(k)
        assert np.array_equal(a, b), "rank %d: %s differs from the one-GPU run (max |diff| %g)" % (
            rank, k, float(np.max(np.abs(a - b))))
    single.close()

    if rank == 0 and w * h <= 512 * 512 and chapter <= 3:
        from oracle import portapi
        ora = portapi.PortSolver(chapter, w, h, 0.1)
        infos_o = [None] * steps
        for i in range(steps):
            ora.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
            infos_o[i] = ora.update(0.005)
        assert [x[:2] for x in infos_m] == [x[:2] for x in infos_o], (infos_m, infos_o)
        for k in "duv":
            a, b = got[k + ".src"], ora.src[k]
            err = float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
            assert err <= 1e-10, (k, err)
    multi.close()
    print("dist ok: rank %d/%d rows [%d,%d) chapter %d %dx%d, %d steps, iterations %s, %d launches" % (
        rank, world, r0, r1, chapter, w, h, steps, [x[1] for x in infos_m], launches), flush=True)


if __name__ == "__main__":
    main()
