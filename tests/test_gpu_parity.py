"""GPU parity tests (run on the B200 box with `-m gpu`): libifl_b200.so, driven through
its C ABI by the Python mirror of FluidSolver, against the oracle (the C restatement of
the reference, oracle/ifl_oracle.c) on identical seeded inputs.

Bars (BASELINE.json north_star):
  * stencil, axpy, MIC(0) factor / solves, advection, assembly, Gauss-Seidel sweeps:
    BIT-EXACT (compared as uint64 words);
  * reduced scalars (dot, and through alpha/beta the PCG iterates): 1e-10 relative in
    double, identical iteration counts.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-10  # north_star tolerance for PCG pressure / velocity / density fields

SIZES = [(32, 32), (33, 35), (64, 64), (100, 70), (70, 131), (257, 129)]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d words differ, first at %d: %r vs %r" %
                             (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]]))


def rel_err(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / scale)


def make_pair(ifl, port, version, w, h, seed=0, plume_steps=0):
    """A device solver and an oracle solver holding identical state."""
    dev = ifl.FluidSolver(w, h, 0.1, version=version)
    ora = port.PortSolver(version, w, h, 0.1)
    rng = np.random.default_rng(seed)
    for k in "duv":
        a = rng.uniform(-1.0, 1.0, ora.src[k].size)
        if k == "d":
            a = np.abs(a)
        ora.src[k][:] = a
        dev.set(k + ".src", a)
    return dev, ora


def sync_vec(dev, ora, name, rng):
    a = rng.uniform(-1.0, 1.0, ora.w * ora.h)
    getattr(ora, name)[:] = a
    dev.set(name, a)


# ------------------------------------------------------------------ assembly -----
@pytest.mark.parametrize("w,h", SIZES)
def test_assembly_bit_exact(ifl, port, w, h):
    dev, ora = make_pair(ifl, port, 3, w, h)
    dev.buildRhs(); ora.buildRhs()
    assert_bits(dev.get("r"), ora.r, "buildRhs")
    dev.buildPressureMatrix(0.005); ora.buildPressureMatrix(0.005)
    for n in ("aDiag", "aPlusX", "aPlusY"):
        assert_bits(dev.get(n), getattr(ora, n), n)
    dev.buildPreconditioner(); ora.buildPreconditioner()
    assert_bits(dev.get("precon"), ora.precon, "buildPreconditioner")
    rng = np.random.default_rng(1)
    sync_vec(dev, ora, "p", rng)
    dev.applyPressure(0.005); ora.applyPressure(0.005)
    assert_bits(dev.get("u.src"), ora.src["u"], "applyPressure u")
    assert_bits(dev.get("v.src"), ora.src["v"], "applyPressure v")
    dev.close()


# ---------------------------------------------------------------- PCG helpers -----
@pytest.mark.parametrize("w,h", SIZES)
def test_pcg_helpers(ifl, port, w, h):
    dev, ora = make_pair(ifl, port, 3, w, h, seed=2)
    for s in (dev, ora):
        s.buildRhs(); s.buildPressureMatrix(0.005); s.buildPreconditioner()
    rng = np.random.default_rng(3)
    for n in ("s", "z", "p"):
        sync_vec(dev, ora, n, rng)
    # matrixVectorProduct v3:315
    dev.matrixVectorProduct("z", "s"); ora.matrixVectorProduct(ora.z, ora.s)
    assert_bits(dev.get("z"), ora.z, "matrixVectorProduct")
    # scaledAdd with every aliasing the reference uses (v3:363, 364, 375)
    dev.scaledAdd("p", "p", "s", 0.37); ora.scaledAdd(ora.p, ora.p, ora.s, 0.37)
    assert_bits(dev.get("p"), ora.p, "scaledAdd p")
    dev.scaledAdd("r", "r", "z", -0.21); ora.scaledAdd(ora.r, ora.r, ora.z, -0.21)
    assert_bits(dev.get("r"), ora.r, "scaledAdd r")
    dev.scaledAdd("s", "z", "s", 1.7); ora.scaledAdd(ora.s, ora.z, ora.s, 1.7)
    assert_bits(dev.get("s"), ora.s, "scaledAdd s")
    # infinityNorm is exact (max is associative); dotProduct is a reordered sum
    assert dev.infinityNorm("r") == ora.infinityNorm(ora.r)
    d_dev, d_ora = dev.dotProduct("z", "r"), ora.dotProduct(ora.z, ora.r)
    norm = float(np.sum(np.abs(ora.z * ora.r)))
    assert abs(d_dev - d_ora) <= 1e-13 * norm
    # applyPreconditioner v3:275 -- exact wavefront execution of both triangular solves
    dev.applyPreconditioner("z", "r"); ora.applyPreconditioner(ora.z, ora.r)
    assert_bits(dev.get("z"), ora.z, "applyPreconditioner")
    dev.close()


def test_preconditioner_signed_zero_rhs(ifl, port):
    """r == -0.0 everywhere except a blob (what buildRhs yields in still regions): the
    boundary handling must not turn -0.0 into +0.0."""
    w, h = 64, 40
    dev, ora = make_pair(ifl, port, 3, w, h, seed=4)
    for s in (dev, ora):
        s.buildPressureMatrix(0.005); s.buildPreconditioner()
    r = np.full(w * h, -0.0)
    r[w * 10 + 5: w * 10 + 20] = np.linspace(-1, 1, 15)
    ora.r[:] = r; dev.set("r", r)
    dev.applyPreconditioner("z", "r"); ora.applyPreconditioner(ora.z, ora.r)
    assert_bits(dev.get("z"), ora.z, "applyPreconditioner(-0.0)")
    dev.close()


# ------------------------------------------------------------------- project -----
@pytest.mark.parametrize("w,h", [(64, 64), (100, 70), (128, 128), (200, 160)])
def test_project_matches_oracle(ifl, port, w, h):
    dev = ifl.FluidSolver(w, h, 0.1, version=3)
    ora = port.PortSolver(3, w, h, 0.1)
    inflow = (0.3, 0.2, 0.3, 0.1, 1.0, 0.4, 3.0)
    dev.addInflow(*inflow); ora.addInflow(*inflow)
    for s in (dev, ora):
        s.buildRhs(); s.buildPressureMatrix(0.005); s.buildPreconditioner()
    st_dev = dev.project(600); st_ora = ora.project(600)
    assert st_dev[0] == st_ora[0] == 0
    assert st_dev[1] == st_ora[1], "iteration count"
    assert rel_err(dev.get("p"), ora.p) <= REL
    assert abs(st_dev[2] - st_ora[2]) <= 1e-9 * max(st_ora[2], 1e-30)
    dev.close()


def test_project_limit_and_early_out(ifl, port):
    w = h = 96
    dev = ifl.FluidSolver(w, h, 0.1, version=3)
    ora = port.PortSolver(3, w, h, 0.1)
    for s in (dev, ora):  # all-zero velocity: |r|inf == 0 -> silent early return v3:355
        s.buildRhs(); s.buildPressureMatrix(0.005); s.buildPreconditioner()
    assert dev.project(600)[0] == ora.project(600)[0] == 2
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    dev.addInflow(*inflow); ora.addInflow(*inflow)
    for s in (dev, ora):
        s.buildRhs()
    r0 = ora.r.copy()
    for limit in (0, 1, 7):  # "Exceeded budget" path v3:379
        ora.r[:] = r0
        dev.set("r", r0)
        st_dev = dev.project(limit); st_ora = ora.project(limit)
        assert st_dev[:2] == st_ora[:2] == (1, limit)
        assert rel_err(dev.get("p"), ora.p) <= REL
        assert rel_err(dev.get("r"), ora.r) <= REL
    assert dev.messages[-1].startswith("Exceeded budget of 7 iterations")
    dev.close()


# -------------------------------------------------------------------- advect -----
@pytest.mark.parametrize("version", [1, 2, 3])
@pytest.mark.parametrize("w,h", [(32, 32), (45, 77), (130, 64)])
def test_advect_bit_exact(ifl, port, version, w, h):
    dev, ora = make_pair(ifl, port, version, w, h, seed=5)
    # velocities large enough to leave the domain in places (exercises the clamps)
    for k in "uv":
        ora.src[k] *= 1000.0 / min(w, h)  # back-trace displacements of up to ~5 cells
        dev.set(k + ".src", ora.src[k])
    for k in "duv":
        dev.advect(k, 0.005); ora.advect(k, 0.005)
    for k in "duv":
        assert_bits(dev.get(k + ".dst"), ora.dst[k], "advect " + k)
    for k in "duv":
        dev.flip(k); ora.flip(k)
    assert_bits(dev.get("u.src"), ora.src["u"], "flip")
    dev.close()


@pytest.mark.parametrize("version", [1, 2])
def test_add_inflow_bit_exact(ifl, port, version):
    w, h = 80, 80
    dev, ora = make_pair(ifl, port, version, w, h, seed=6)
    for s in (dev, ora):
        s.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        s.addInflow(-0.1, 0.9, 0.5, 0.3, 0.5, -2.0, 0.25)  # clipped by the domain
    for k in "duv":
        assert_bits(dev.get(k + ".src"), ora.src[k], "addInflow " + k)
    dev.close()


@pytest.mark.parametrize("version", [2, 3])
def test_add_inflow_wraps_rows_on_tall_grids(ifl, port, version):
    """w < h: the reference clamps the x loop with _h (v2:195), so a wide rectangle runs past the row and
    its dense index x + y*_w lands in the following rows.  The device reproduces that wrap."""
    w, h = 40, 100
    dev, ora = make_pair(ifl, port, version, w, h, seed=8)
    for s in (dev, ora):
        s.addInflow(0.5, 0.3, 1.0, 0.4, 1.0, -0.5, 2.0)   # ix1 = 59 > w: 19 cells of every row wrap into the next
        s.addInflow(0.1, 1.0, 2.4, 0.2, 0.7, 0.3, -1.0)   # ix1 = 99: wraps twice
    for k in "duv":
        assert_bits(dev.get(k + ".src"), ora.src[k], "addInflow (wrapped) " + k)
    dev.close()


# -------------------------------------------------------------- Gauss-Seidel -----
@pytest.mark.parametrize("w,h", [(32, 32), (33, 35), (64, 96), (100, 70)])
def test_gauss_seidel_sweeps_bit_exact(ifl, port, w, h):
    dev, ora = make_pair(ifl, port, 2, w, h, seed=7)
    dev.buildRhs(); ora.buildRhs()
    assert_bits(dev.get("r"), ora.r, "buildRhs")
    for limit in (1, 3, 10):  # warm-started: p carries over (v2:308)
        st_dev = dev.project(limit, 0.005); st_ora = ora.project(limit, 0.005)
        assert_bits(dev.get("p"), ora.p, "GS p after %d more sweeps" % limit)
        assert st_dev[:2] == st_ora[:2]
        assert st_dev[2] == st_ora[2]
    dev.close()


def test_gauss_seidel_converges_same_iteration(ifl, port):
    w = h = 32
    dev = ifl.FluidSolver(w, h, 0.1, version=2)
    ora = port.PortSolver(2, w, h, 0.1)
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    dev.addInflow(*inflow); ora.addInflow(*inflow)
    dev.buildRhs(); ora.buildRhs()
    st_dev = dev.project(600, 0.005); st_ora = ora.project(600, 0.005)
    assert st_dev == st_ora
    assert st_ora[0] == 0, "expected the 32^2 solve to converge inside the budget"
    assert_bits(dev.get("p"), ora.p, "GS converged p")
    assert dev.messages[-1] == "Exiting solver after %d iterations, maximum change is %f" % (st_ora[1], st_ora[2])
    dev.close()


# ---------------------------------------------------------------- trajectories -----
@pytest.mark.parametrize("version,w,h,steps", [(3, 128, 128, 10), (3, 96, 160, 5), (2, 64, 64, 4), (1, 64, 64, 4)])
def test_update_trajectory(ifl, port, version, w, h, steps):
    dev = ifl.FluidSolver(w, h, 0.1, version=version)
    ora = port.PortSolver(version, w, h, 0.1)
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    for i in range(steps):
        dev.addInflow(*inflow); ora.addInflow(*inflow)
        st_dev = dev.update(0.005); st_ora = ora.update(0.005)
        assert st_dev[:2] == st_ora[:2], "step %d solver status" % i
    for k in "duv":
        if version < 3:  # Gauss-Seidel chapters are bit-exact end to end
            assert_bits(dev.get(k + ".src"), ora.src[k], k)
        else:
            assert rel_err(dev.get(k + ".src"), ora.src[k]) <= REL, k
    dev.close()


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_update_matches_reference_golden(ifl, path):
    """Device trajectory against vectors produced by the UNMODIFIED reference."""
    g = np.load(path)
    ver = int(g["version"])
    dev = ifl.FluidSolver(int(g["w"]), int(g["h"]), float(g["density"]), version=ver)
    iters = []
    for _ in range(int(g["steps"])):
        dev.addInflow(*g["inflow"])
        iters.append(dev.update(float(g["timestep"]))[1])
    assert iters == list(g["iters"])
    for k in "duv":
        if ver < 3:
            assert_bits(dev.get(k + ".src"), g[k], k)
        else:
            assert rel_err(dev.get(k + ".src"), g[k]) <= REL, k
    assert rel_err(dev.get("p"), g["p"]) <= REL
    dev.close()


def test_update_host_equals_update(ifl):
    """The end-to-end entry point (host buffers in/out) is the same step."""
    w = h = 96
    a = ifl.FluidSolver(w, h, 0.1, version=3)
    b = ifl.FluidSolver(w, h, 0.1, version=3)
    a.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    d, u, v = a.get("d.src"), a.get("u.src"), a.get("v.src")
    a.update(0.005)
    b.update_host(0.005, d, u, v)
    for k, arr in (("d", d), ("u", u), ("v", v)):
        assert_bits(a.get(k + ".src"), arr, k)
    a.close(); b.close()


# ------------------------------------------------------- full-size properties -----
def test_preconditioner_roundtrip_large(ifl):
    """Size-independent check of the wavefront machinery at 2048^2 (64 strips x 64
    blocks): M = L L^T, so multiplying the solve's result back must return the rhs.
    The inverse relations are evaluated with numpy vector ops (no sequential loop)."""
    n = 2048
    dev = ifl.FluidSolver(n, n, 0.1, version=3)
    dev.buildPressureMatrix(0.005); dev.buildPreconditioner()
    rng = np.random.default_rng(8)
    r = rng.uniform(-1, 1, n * n)
    dev.set("r", r)
    dev.applyPreconditioner("z", "r")
    z = dev.get("z").reshape(n, n)
    pre = dev.get("precon").reshape(n, n)
    cx = dev.get("aPlusX").reshape(n, n) * pre
    cy = dev.get("aPlusY").reshape(n, n) * pre
    # backward: z = (y - cx*z[x+1] - cy*z[y+1]) * pre   ->  y
    y = z / pre
    y[:, :-1] += cx[:, :-1] * z[:, 1:]
    y[:-1, :] += cy[:-1, :] * z[1:, :]
    # forward: y = (r - cx[x-1]*y[x-1] - cy[y-1]*y[y-1]) * pre  ->  r
    rr = y / pre
    rr[:, 1:] += cx[:, :-1] * y[:, :-1]
    rr[1:, :] += cy[:-1, :] * y[:-1, :]
    assert rel_err(rr.ravel(), r) <= 1e-9
    dev.close()


def test_matvec_symmetry_large(ifl):
    """x.(A y) == y.(A x) at 2048^2 through the device kernels (A is symmetric)."""
    n = 2048
    dev = ifl.FluidSolver(n, n, 0.1, version=3)
    dev.buildPressureMatrix(0.005)
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, n * n); y = rng.uniform(-1, 1, n * n)
    dev.set("s", x); dev.matrixVectorProduct("z", "s"); dev.set("r", y)
    a = dev.dotProduct("z", "r")
    dev.set("s", y); dev.matrixVectorProduct("z", "s"); dev.set("r", x)
    b = dev.dotProduct("z", "r")
    assert abs(a - b) <= 1e-9 * max(abs(a), abs(b))
    dev.close()


@pytest.mark.parametrize("version", [3, 2])
@pytest.mark.parametrize("w,h", [(2, 2), (3, 3), (5, 40), (40, 5), (31, 33), (33, 31), (64, 20), (17, 257), (300, 34)])
def test_update_edge_sizes(ifl, port, version, w, h):
    """Degenerate and ragged grids (the reference accepts any w, h >= 2): fewer rows than one
    strip, one column block, sizes just off the 32-cell tiles, a cluster with padding CTAs."""
    dev = ifl.FluidSolver(w, h, 0.1, version=version)
    ora = port.PortSolver(version, w, h, 0.1)
    for _ in range(3):
        dev.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        ora.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        assert dev.update(0.005)[:2] == ora.update(0.005)[:2]
    for k in "duv":
        a, b = dev.get(k + ".src"), ora.src[k]
        err = float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
        assert err <= 1e-10, (k, err)  # PCG results: <= 1e-10 relative (north star)
    dev.close()


# ---- every sweep engine and cluster size does the same arithmetic ---------------------------------
# The product runs the two-row engine in clusters of 16 / 8; the one-row engine (IFL_TRI=0), the staircase
# engine (IFL_TRI=2), plain clusters of 8 and the serial order of k_axpy2_norm and the
# forward sweep (IFL_OVERLAP_AXPY=0) are alternatives selected when a solver is created.
ENGINE_ENVS = [{"IFL_TRI": "0"}, {"IFL_TRI": "2"}, {"IFL_TRI": "2", "IFL_OVERLAP_AXPY": "0"}, {"IFL_TRI_CLUSTER16": "0"},
               {"IFL_OVERLAP_AXPY": "0"}, {"IFL_OVERLAP_AXPY": "1"}, {"IFL_SWEEP_CLUSTER": "1"}]


@pytest.mark.gpu
@pytest.mark.parametrize("env", ENGINE_ENVS, ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
@pytest.mark.parametrize("w,h", [(257, 130), (64, 700), (1024, 1024)])
def test_engine_variants_bit_exact(ifl, port, monkeypatch, env, w, h):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    dev, ora = make_pair(ifl, port, 3, w, h, seed=11)
    for s in (dev, ora):
        s.buildRhs(); s.buildPressureMatrix(0.005); s.buildPreconditioner()
    assert_bits(dev.get("precon"), ora.precon, "buildPreconditioner")
    dev.applyPreconditioner("z", "r"); ora.applyPreconditioner(ora.z, ora.r)
    assert_bits(dev.get("z"), ora.z, "applyPreconditioner")
    sd, so = dev.project(40), ora.project(40)
    assert sd[:2] == so[:2], (sd, so)
    assert rel_err(dev.get("p"), ora.p) <= REL
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"IFL_TRI": "2"}, {"IFL_TRI": "0"}], ids=["staircase", "one-row"])
def test_engine_variants_with_solids(ifl, monkeypatch, env):
    """Masked form (v5:746-780) of the alternative engines against the unmodified reference."""
    from oracle import refapi
    import math
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    w, h = 200, 136
    bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0), ifl.SolidSphere(0.2, 0.3, 0.2, 0.0, 0.0, 0.0, 0.0)]
    dev = ifl.FluidSolver(w, h, 0.1, version=5, bodies=bodies)
    ref = refapi.Ref(5, w, h, [0.1], [b.as_row() for b in bodies])
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    dev.addInflow(*inflow); ref.call("addInflow", *inflow)
    for q in "duv":
        dev.fillSolidFields(q); ref.call(q + ".fillSolidFields")
    dev.setBoundaryCondition(); ref.call("setBoundaryCondition")
    dev.buildRhs(); ref.call("buildRhs")
    dev.buildPressureMatrix(0.005); ref.call("buildPressureMatrix", 0.005)
    dev.buildPreconditioner(); ref.call("buildPreconditioner")
    assert_bits(dev.get("precon"), ref.buf("precon"), "precon")
    dev.applyPreconditioner("z", "r"); ref.call("applyPreconditioner", 2, 0)
    assert_bits(dev.get("z"), ref.buf("z"), "applyPreconditioner (masked)")
    dev.close(); ref.close()


@pytest.mark.gpu
def test_matvec_uses_an_uploaded_matrix(ifl, port):
    """Chapter 3's k_matvec evaluates the matrix buildPressureMatrix wrote from the cell position; a matrix
    the caller uploads afterwards must be READ again, and the next buildPressureMatrix re-enables the shortcut."""
    w, h = 300, 140
    dev, ora = make_pair(ifl, port, 3, w, h, seed=21)
    rng = np.random.default_rng(22)
    dev.buildPressureMatrix(0.005); ora.buildPressureMatrix(0.005)
    sync_vec(dev, ora, "s", rng)
    dev.matrixVectorProduct("z", "s"); ora.matrixVectorProduct(ora.z, ora.s)
    assert_bits(dev.get("z"), ora.z, "matrixVectorProduct (position-only matrix)")
    for name in ("aDiag", "aPlusX", "aPlusY"):
        a = getattr(ora, name) * rng.uniform(0.5, 1.5, w * h)
        getattr(ora, name)[:] = a
        dev.set(name, a)
    dev.matrixVectorProduct("z", "s"); ora.matrixVectorProduct(ora.z, ora.s)
    assert_bits(dev.get("z"), ora.z, "matrixVectorProduct (uploaded matrix)")
    dev.buildPressureMatrix(0.0025); ora.buildPressureMatrix(0.0025)
    dev.matrixVectorProduct("z", "s"); ora.matrixVectorProduct(ora.z, ora.s)
    assert_bits(dev.get("z"), ora.z, "matrixVectorProduct (rebuilt matrix)")
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [1, 2, 3, 4, 9, 400])
def test_solver_vectors_after_project(ifl, port, limit):
    """r, z, s and p as the reference leaves them when the budget runs out after `limit` iterations or the solve
    converges (limit 400) -- chapter 3 keeps s in a ping-pong pair whose halves are swapped once per fused launch
    (DESIGN 3.2), an odd and an even number of times here."""
    w, h = 200, 136
    dev = ifl.FluidSolver(w, h, 0.1, version=3)
    ora = port.PortSolver(3, w, h, 0.1)
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    for s in (dev, ora):
        s.addInflow(*inflow)
        s.buildRhs(); s.buildPressureMatrix(0.005); s.buildPreconditioner()
    sd, so = dev.project(limit), ora.project(limit)
    assert sd[:2] == so[:2], (sd, so)
    converged = sd[0] == 0
    assert converged == (limit == 400), sd
    # (a converged reference solve returns with A*s still in _z, v3:361-368: the device keeps that product in `q`)
    for name in ("r", "s", "p") if converged else ("r", "z", "s", "p"):
        assert rel_err(dev.get(name), getattr(ora, name)) <= 1e-9, (limit, name)
    # and the granular entry points keep working on the buffer that now is `s`
    dev.matrixVectorProduct("z", "s"); ora.matrixVectorProduct(ora.z, ora.s)
    assert rel_err(dev.get("z"), ora.z) <= 1e-9
    dev.close()
