"""Parity at the sizes that are benchmarked ("parity at the sizes you benchmark").

Every kernel class of the chapter-3 hot path is compared with the UNMODIFIED reference
(oracle/_ref/libref_v3.so) at 4096^2 -- 64 strips of the triangular-solve engine in 8 clusters, 128
strips of the factorisation engine in 16 -- and the regime "more strips than SMs" (strip tickets are
handed out in several waves) is covered by tall narrow grids through the oracle port.

Bars: bit-exact for buildPreconditioner, applyPreconditioner, matrixVectorProduct, advect,
applyPressure, Gauss-Seidel sweeps; 1e-10 relative for project(limit=5) (the only difference
inside it is the summation order of the dot products).
"""
import numpy as np
import pytest

from oracle import refapi

pytestmark = pytest.mark.gpu

REL = 1e-10
N = 4096


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d words differ, first at %d: %r vs %r" %
                             (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]]))


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def pair(ifl):
    """Device solver and reference solver at 4096^2 with identical plume-like state."""
    if not refapi.available(3):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    dev = ifl.FluidSolver(N, N, 0.1, version=3)
    ref = refapi.Ref(3, N, N, [0.1])
    rng = np.random.default_rng(7)
    for k in "duv":
        a = rng.uniform(-1.0, 1.0, ref.buf(k + ".src").size)
        if k == "d":
            a = np.abs(a)
        ref.buf(k + ".src")[:] = a
        dev.set(k + ".src", a)
    yield dev, ref
    dev.close()
    ref.close()


def test_4096_assembly_and_factorisation(pair):
    dev, ref = pair
    dev.buildRhs(); ref.call("buildRhs")
    assert_bits(dev.get("r"), ref.buf("r"), "buildRhs")
    dev.buildPressureMatrix(0.005); ref.call("buildPressureMatrix", 0.005)
    for n in ("aDiag", "aPlusX", "aPlusY"):
        assert_bits(dev.get(n), ref.buf(n), n)
    dev.buildPreconditioner(); ref.call("buildPreconditioner")
    assert_bits(dev.get("precon"), ref.buf("precon"), "buildPreconditioner")


def test_4096_apply_preconditioner(pair):
    dev, ref = pair  # matrix + factorisation from the previous test
    dev.applyPreconditioner("z", "r"); ref.call("applyPreconditioner", 2, 0)
    assert_bits(dev.get("z"), ref.buf("z"), "applyPreconditioner")


def test_4096_matrix_vector_product(pair):
    dev, ref = pair
    rng = np.random.default_rng(11)
    s = rng.uniform(-1.0, 1.0, N * N)
    ref.buf("s")[:] = s
    dev.set("s", s)
    dev.matrixVectorProduct("z", "s"); ref.call("matrixVectorProduct", 2, 3)
    assert_bits(dev.get("z"), ref.buf("z"), "matrixVectorProduct")


def test_4096_project_5_iterations(pair):
    dev, ref = pair
    dev.buildRhs(); ref.call("buildRhs")
    st = dev.project(5)
    ref.log()
    ref.call("project", 5)
    log = ref.log()
    assert "Exceeded budget of 5 iterations" in log
    assert st[0] == 1 and st[1] == 5, st  # status 1 = IFL_SOLVE_EXCEEDED
    # the line the drop-in prints is the reference's, up to the 6 printed digits of the residual
    assert dev.messages[-1].split(",")[0] == log.strip().splitlines()[-1].split(",")[0]
    assert rel_err(dev.get("p"), ref.buf("p")) <= REL
    assert rel_err(dev.get("r"), ref.buf("r")) <= REL


def test_4096_apply_pressure_and_advect(pair):
    dev, ref = pair
    rng = np.random.default_rng(13)
    p = rng.uniform(-1.0, 1.0, N * N)
    ref.buf("p")[:] = p
    dev.set("p", p)
    dev.applyPressure(0.005); ref.call("applyPressure", 0.005)
    assert_bits(dev.get("u.src"), ref.buf("u.src"), "applyPressure u")
    assert_bits(dev.get("v.src"), ref.buf("v.src"), "applyPressure v")
    # one field is enough at this size (the reference needs ~4 s per field): d samples the same u, v
    dev.advect("d", 0.005); ref.call("d.advect", 0.005)
    assert_bits(dev.get("d.dst"), ref.buf("d.dst"), "advect d")


# ---- more strips than SMs: tickets are handed out in several waves -------------------------------
# On these 94:1 .. 300:1 grids the solve needs thousands of iterations, so update() stops on the
# reference's cap of 600 with a residual that is still large, and 600 iterations of CG amplify ANY
# rounding difference far beyond 1e-10.  The kernels are therefore pinned bit-exactly / at 1e-10 by a
# short solve, and the whole capped steps are held against a MEASURED envelope: the unmodified reference's
# own kernels run again with nothing but the summation order of the dot products changed -- the liberty the
# device takes -- (tests/envelope.py: update_reordered).
from envelope import Envelope, update_reordered  # noqa: E402


@pytest.mark.parametrize("w,h", [(64, 6016), (96, 16384), (40, 12000)])
def test_tall_grid_update_vs_oracle(ifl, port, w, h):
    """64 x 6016 = 188 strips of the one-row engine / 94 of the two-row engine; 96 x 16384 = 512 / 256
    (the multi-wave ticket path of both engines)."""
    dev = ifl.FluidSolver(w, h, 0.1, version=3)
    ora = port.PortSolver(3, w, h, 0.1)
    twin = refapi.Ref(3, w, h, [0.1])  # the unmodified reference, its reductions re-ordered (tests/envelope.py)
    inflow = (0.2, 0.2, 0.3, 0.5, 1.0, 0.0, 3.0)
    for s in (dev, ora):
        s.addInflow(*inflow)
    twin.call("addInflow", *inflow)
    # short solve on the stamped plume: every kernel of the loop, tight bar
    dev.buildRhs(); ora.buildRhs()
    dev.buildPressureMatrix(0.005); ora.buildPressureMatrix(0.005)
    dev.buildPreconditioner(); ora.buildPreconditioner()
    assert_bits(dev.get("precon"), ora.precon, "precon")
    sd = dev.project(8)
    so = ora.project(8)
    assert sd[:2] == so[:2], (sd, so)
    assert rel_err(dev.get("p"), ora.p) <= REL
    # whole capped steps against the reference's own one-ulp envelope
    envelope = Envelope(floor=REL)
    for step in range(2):
        sd = dev.update(0.005)
        so = ora.update(0.005)
        update_reordered(twin, 0.005)
        assert sd[:2] == so[:2], (step, sd, so)
        for k in "duv":
            e = rel_err(dev.get(k + ".src"), ora.src[k])
            env = rel_err(twin.buf(k + ".src"), ora.src[k])
            print("tall %dx%d step %d %s: device vs reference %.2e, reference vs re-ordered reference %.2e" % (w, h, step, k, e, env))
            envelope.check(e, env, (step, k))
        for s in (dev, ora):
            s.addInflow(*inflow)
        twin.call("addInflow", *inflow)
    dev.close()
    twin.close()


@pytest.mark.parametrize("w,h", [(64, 6016)])
def test_tall_grid_preconditioner_bit_exact(ifl, port, w, h):
    dev = ifl.FluidSolver(w, h, 0.1, version=3)
    ora = port.PortSolver(3, w, h, 0.1)
    rng = np.random.default_rng(3)
    dev.buildPressureMatrix(0.005); ora.buildPressureMatrix(0.005)
    dev.buildPreconditioner(); ora.buildPreconditioner()
    assert_bits(dev.get("precon"), ora.precon, "precon")
    a = rng.uniform(-1.0, 1.0, w * h)
    ora.r[:] = a
    dev.set("r", a)
    dev.applyPreconditioner("z", "r"); ora.applyPreconditioner(ora.z, ora.r)
    assert_bits(dev.get("z"), ora.z, "applyPreconditioner")
    dev.close()


def test_gauss_seidel_2048(ifl, port):
    """Chapter 2 at config-2 size: two lexicographic sweeps, bit-exact incl. the max |dp| the stop test sees."""
    n = 2048
    dev = ifl.FluidSolver(n, n, 0.1, version=2)
    ora = port.PortSolver(2, n, n, 0.1)
    rng = np.random.default_rng(5)
    for k in "uv":
        a = rng.uniform(-1.0, 1.0, ora.src[k].size)
        ora.src[k][:] = a
        dev.set(k + ".src", a)
    dev.buildRhs(); ora.buildRhs()
    sd = dev.project(2, 0.005)
    so = ora.project(2, 0.005)
    assert sd[:2] == so[:2], (sd, so)
    assert_bits(dev.get("p"), ora.p, "Gauss-Seidel p after 2 sweeps")
    dev.close()
