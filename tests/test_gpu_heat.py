"""GPU parity tests for chapters 6-7 (heat diffusion, buoyancy, variable density) against
the UNMODIFIED reference (oracle/_ref/libref_v6.so, libref_v7.so)."""
import math
import re

import numpy as np
import pytest

from oracle import refapi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refapi.available(7), reason="oracle/_ref not built (needs /root/reference)")]

RHO_AIR, RHO_SOOT, DIFFUSION = 0.1, 1.0, 0.01  # v7:1091-1093 (v6 ships rhoSoot 0.1)
SOLID_REL = 2e-6  # reference noise floor with solid bodies, see tests/test_gpu_solids.py


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d differ, first at %d: %r vs %r" %
                             (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]]))


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def make_pair(ifl, version, w, h):
    bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0)]  # v7:1099
    dev = ifl.FluidSolver(w, h, RHO_AIR, version=version, bodies=bodies, rho_soot=RHO_SOOT, diffusion=DIFFUSION)
    ref = refapi.Ref(version, w, h, [RHO_AIR, RHO_SOOT, DIFFUSION], [b.as_row() for b in bodies])
    return dev, ref, bodies


def seed(dev, ref, rng):
    for k in "duv":
        a = rng.uniform(-1.0, 1.0, ref.buf(k + ".src").size)
        if k == "d":
            a = np.abs(a)
        ref.buf(k + ".src")[:] = a
        dev.set(k + ".src", a)
    t = 294.0 + rng.uniform(0.0, 300.0, ref.buf("t.src").size)
    ref.buf("t.src")[:] = t
    dev.set("t.src", t)


@pytest.mark.parametrize("version", [6, 7])
def test_heat_density_stages_bit_exact(ifl, version):
    w = h = 96
    dev, ref, _ = make_pair(ifl, version, w, h)
    assert dev.ambientT() == ref.call("ambientT") == 294.0
    assert_bits(dev.get("t.src"), ref.buf("t.src"), "initial temperature")
    seed(dev, ref, np.random.default_rng(0))
    for k in "dtuv":
        dev.fillSolidFields(k); ref.call(k + ".fillSolidFields")
    # heat system
    dev.buildHeatDiffusionMatrix(0.005); ref.call("buildHeatDiffusionMatrix", 0.005)
    for n in ("aDiag", "aPlusX", "aPlusY"):
        assert_bits(dev.get(n), ref.buf(n), "heat " + n)
    dev.buildPreconditioner(); ref.call("buildPreconditioner")
    fluid = ref.buf("d.cell") == 0
    assert_bits(dev.get("precon")[fluid], ref.buf("precon")[fluid], "heat precon")
    # masked vector helpers (v6:781-826): non-fluid entries must stay untouched
    rng = np.random.default_rng(1)
    for n in ("r", "s", "z", "p"):
        a = rng.uniform(-1, 1, w * h)
        ref.buf(n)[:] = a; dev.set(n, a)
    dev.scaledAdd("p", "p", "s", 0.3); ref.call("scaledAdd", 1, 1, 3, 0.3)
    assert_bits(dev.get("p"), ref.buf("p"), "masked scaledAdd")
    assert dev.infinityNorm("r") == ref.call("infinityNorm", 0)
    dd, dr = dev.dotProduct("z", "r"), ref.call("dotProduct", 2, 0)
    assert abs(dd - dr) <= 1e-13 * float(np.sum(np.abs(ref.buf("z") * ref.buf("r"))))
    # buoyancy, densities, pressure system
    dev.addBuoyancy(0.005); ref.call("addBuoyancy", 0.005)
    assert_bits(dev.get("v.src"), ref.buf("v.src"), "addBuoyancy")
    dev.setBoundaryCondition(); ref.call("setBoundaryCondition")
    dev.buildRhs(); ref.call("buildRhs")
    assert_bits(dev.get("r"), ref.buf("r"), "buildRhs")
    if version >= 7:
        dev.computeDensities(); ref.call("computeDensities")
        assert_bits(dev.get("uDensity"), ref.buf("uDensity"), "uDensity")
        assert_bits(dev.get("vDensity"), ref.buf("vDensity"), "vDensity")
    dev.buildPressureMatrix(0.005); ref.call("buildPressureMatrix", 0.005)
    for n in ("aDiag", "aPlusX", "aPlusY"):
        assert_bits(dev.get(n), ref.buf(n), "pressure " + n)
    p = rng.uniform(-1, 1, w * h)
    dev.set("p", p); ref.buf("p")[:] = p
    dev.applyPressure(0.005); ref.call("applyPressure", 0.005)
    assert_bits(dev.get("u.src"), ref.buf("u.src"), "applyPressure u")
    assert_bits(dev.get("v.src"), ref.buf("v.src"), "applyPressure v")
    dev.close(); ref.close()


@pytest.mark.parametrize("version,w,h,steps", [(7, 128, 128, 5), (6, 96, 96, 4)])
def test_update_heat_stepwise(ifl, version, w, h, steps):
    """update() (v7:995-1034) one step at a time from the reference's state."""
    dev, ref, bodies = make_pair(ifl, version, w, h)
    tamb = dev.ambientT()
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, tamb, 0.0, 0.0)  # v7:1112
    for i in range(steps):
        for k in "dtuv":
            dev.set(k + ".src", ref.buf(k + ".src"))
            dev.set(k + ".dst", ref.buf(k + ".dst"))
        dev.addInflow(*inflow); ref.call("addInflow", *inflow)
        st = dev.update(0.005)
        ref.call("update", 0.005)
        its = [int(x) for x in re.findall(r"(?:after|of) (\d+) iterations", ref.log())]
        assert len(its) == 2, its
        assert abs(dev.last_heat[1] - its[0]) <= 1, (dev.last_heat, its)
        assert abs(st[1] - its[1]) <= max(2, 0.06 * its[1]), (st, its)
        for k in "dtuv":
            assert rel_err(dev.get(k + ".src"), ref.buf(k + ".src")) <= SOLID_REL, (i, k)
    dev.close(); ref.close()


# ---- whole trajectories without re-synchronisation, against a MEASURED envelope (see test_gpu_solids.py) ----
from envelope import Envelope, update_reordered  # noqa: E402


@pytest.mark.parametrize("version,rho_soot,steps", [(7, 1.0, 10), (6, 0.1, 10)])
def test_heat_trajectory_within_measured_envelope(ifl, version, rho_soot, steps):
    """10 steps from the constructors on with the shipped constants (v7:1084-1129 / v6:1058-1100): device,
    reference, and the reference's own kernels with re-ordered reductions (tests/envelope.py).
    The device has to stay within ENV_FACTOR times what the twin drifts, iteration counts within its spread."""
    w = h = 128
    bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0)]
    rows = [b.as_row() for b in bodies]
    dev = ifl.FluidSolver(w, h, RHO_AIR, version=version, bodies=bodies, rho_soot=rho_soot, diffusion=DIFFUSION)
    ref = refapi.Ref(version, w, h, [RHO_AIR, rho_soot, DIFFUSION], rows)
    twin = refapi.Ref(version, w, h, [RHO_AIR, rho_soot, DIFFUSION], rows, fresh_copy=True)  # own copy of the library: own log
    tamb = dev.ambientT()
    inflow = (0.45, 0.2, 0.1, 0.05, 1.0, tamb + (300.0 if version == 6 else 0.0), 0.0, 0.0)  # v7:1112 / v6:1086
    envelope = Envelope(floor=steps * 1e-10)  # `steps` chained pairs of solves
    for step in range(steps):
        dev.addInflow(*inflow); ref.call("addInflow", *inflow); twin.call("addInflow", *inflow)
        st = dev.update(0.005)
        ref.call("update", 0.005)
        itw = update_reordered(twin, 0.005)
        its = [int(x) for x in re.findall(r"(?:after|of) (\d+) iterations", ref.log())]
        assert len(its) == 2 and len(itw) == 2, (its, itw)
        Envelope.check_iterations(dev.last_heat[1], its[0], itw[0], (step, "heat"))
        Envelope.check_iterations(st[1], its[1], itw[1], (step, "pressure"))
        for k in "dtuv":
            e = rel_err(dev.get(k + ".src"), ref.buf(k + ".src"))
            env = rel_err(twin.buf(k + ".src"), ref.buf(k + ".src"))
            print("chapter %d step %d %s: device vs reference %.2e, reference vs re-ordered reference %.2e" % (version, step, k, e, env))
            envelope.check(e, env, (step, k))
    print("chapter %d: worst device deviation %.2e at a reference envelope of %.2e" % ((version,) + envelope.worst))
    dev.close(); ref.close(); twin.close()
