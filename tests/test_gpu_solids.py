"""GPU parity tests for chapters 4-5 (solid bodies, curved boundaries): the device path
against the UNMODIFIED reference (oracle/_ref/libref_v4.so, libref_v5.so -- prebuilt
where /root/reference exists, they travel with the repo snapshot).

Bit-exact: fillSolidFields (cell, body, volume, normals), setBoundaryCondition, buildRhs,
buildPressureMatrix, buildPreconditioner, applyPreconditioner, applyPressure,
extrapolate, advect.  PCG results: 1e-10 relative, identical iteration counts.
"""
import math

import numpy as np
import pytest

from oracle import refapi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refapi.available(5), reason="oracle/_ref not built (needs /root/reference)")]

REL = 1e-10


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    same = np.array_equal(bits(a), bits(b)) if a.dtype == np.float64 else np.array_equal(a, b)
    if not same:
        bad = np.flatnonzero((bits(a) != bits(b)) if a.dtype == np.float64 else (a != b))
        raise AssertionError("%s: %d of %d differ, first at %d: %r vs %r" %
                             (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]]))


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def make_bodies(ifl, moving=True):
    """rotating box, sphere, counter-rotating small box."""
    w = 1.0 if moving else 0.0
    return [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 1.0 * w),
            ifl.SolidSphere(0.15, 0.3, 0.15, 0.0, 0.0, 0.0, 0.0),
            ifl.SolidBox(0.85, 0.2, 0.2, 0.1, 0.0, 0.0, 0.0, -2.0 * w)]


def make_pair(ifl, version, w, h, moving=True):
    bodies = make_bodies(ifl, moving)
    dev = ifl.FluidSolver(w, h, 0.1, version=version, bodies=bodies)
    ref = refapi.Ref(version, w, h, [0.1], [b.as_row() for b in bodies])
    return dev, ref, bodies


def seed_fields(dev, ref, seed=0):
    rng = np.random.default_rng(seed)
    for k in "duv":
        a = rng.uniform(-1.0, 1.0, ref.buf(k + ".src").size)
        ref.buf(k + ".src")[:] = a
        dev.set(k + ".src", a)


@pytest.mark.parametrize("version", [4, 5])
@pytest.mark.parametrize("w,h", [(64, 64), (100, 100), (130, 130)])
def test_fill_solid_fields_bit_exact(ifl, version, w, h):
    dev, ref, _ = make_pair(ifl, version, w, h)
    for k in "duv":
        dev.fillSolidFields(k)
        ref.call(k + ".fillSolidFields")
        for which in ("cell", "body", "normalX", "normalY") + (("volume", "phi") if version >= 5 else ()):
            assert_bits(dev.get_aux(k, which), ref.buf("%s.%s" % (k, which)), "%s.%s" % (k, which))
    dev.close(); ref.close()


@pytest.mark.parametrize("version", [4, 5])
def test_assembly_and_preconditioner_bit_exact(ifl, version):
    w = h = 96
    dev, ref, _ = make_pair(ifl, version, w, h)
    seed_fields(dev, ref, 1)
    for k in "duv":
        dev.fillSolidFields(k); ref.call(k + ".fillSolidFields")
    dev.setBoundaryCondition(); ref.call("setBoundaryCondition")
    assert_bits(dev.get("u.src"), ref.buf("u.src"), "setBoundaryCondition u")
    assert_bits(dev.get("v.src"), ref.buf("v.src"), "setBoundaryCondition v")
    dev.buildRhs(); ref.call("buildRhs")
    assert_bits(dev.get("r"), ref.buf("r"), "buildRhs")
    dev.buildPressureMatrix(0.005); ref.call("buildPressureMatrix", 0.005)
    for n in ("aDiag", "aPlusX", "aPlusY"):
        assert_bits(dev.get(n), ref.buf(n), n)
    dev.buildPreconditioner(); ref.call("buildPreconditioner")
    assert_bits(dev.get("precon"), ref.buf("precon"), "buildPreconditioner")
    # applyPreconditioner(z, r): fluid cells bit-exact; non-fluid cells keep their old z
    dev.applyPreconditioner("z", "r"); ref.call("applyPreconditioner", 2, 0)
    fluid = ref.buf("d.cell") == 0
    assert_bits(dev.get("z")[fluid], ref.buf("z")[fluid], "applyPreconditioner")
    rng = np.random.default_rng(2)
    p = rng.uniform(-1, 1, w * h)
    dev.set("p", p); ref.buf("p")[:] = p
    dev.applyPressure(0.005); ref.call("applyPressure", 0.005)
    assert_bits(dev.get("u.src"), ref.buf("u.src"), "applyPressure u")
    assert_bits(dev.get("v.src"), ref.buf("v.src"), "applyPressure v")
    dev.close(); ref.close()


@pytest.mark.parametrize("version", [4, 5])
def test_extrapolate_and_advect_bit_exact(ifl, version):
    w = h = 112
    dev, ref, _ = make_pair(ifl, version, w, h)
    seed_fields(dev, ref, 3)
    for k in "uv":  # a few cells of back-trace displacement
        a = ref.buf(k + ".src") * (600.0 / w)
        ref.buf(k + ".src")[:] = a
        dev.set(k + ".src", a)
    for k in "duv":
        dev.fillSolidFields(k); ref.call(k + ".fillSolidFields")
    for k in "duv":
        dev.extrapolate(k); ref.call(k + ".extrapolate")
        assert_bits(dev.get(k + ".src"), ref.buf(k + ".src"), "extrapolate " + k)
    for k in "duv":
        dev.advect(k, 0.005); ref.call(k + ".advect", 0.005)
    for k in "duv":
        cell = ref.buf(k + ".cell") == 0
        assert_bits(dev.get(k + ".dst")[cell], ref.buf(k + ".dst")[cell], "advect " + k)
        # non-fluid cells are not written: both sides still hold zeros
        assert_bits(dev.get(k + ".dst")[~cell], ref.buf(k + ".dst")[~cell], "advect (untouched) " + k)
    dev.close(); ref.close()


# ---- whole update() with solid bodies -----------------------------------------------------
# With solid bodies the all-Neumann pressure system is singular and its right-hand side is
# only approximately compatible, so the reference's PCG is *itself* extremely sensitive to
# rounding: flipping the last bit of ONE rhs entry changes the unmodified reference's own
# converged pressure by ~1e-9..1e-8 relative (test_reference_pcg_sensitivity below pins
# that number).  Every stage that feeds the solve is bit-identical on the device (tests
# above); the only difference inside the solve is the summation order of the two dot
# products, which perturbs alpha/beta in the last bits and is amplified the same way.  The
# trajectory bar for these chapters is therefore the reference's own noise floor, not the
# 1e-10 that holds without solids (tests/test_gpu_parity.py).
SOLID_REL = 2e-6   # per step from identical state; observed 4e-9 .. 5e-7
ITER_SLACK = 0.06  # |iterations - reference| / reference


def one_step_from_reference_state(dev, ref, bodies, inflow):
    """Copies the reference's current d,u,v into the device, steps both once, returns
    (device status, reference iteration count)."""
    import re
    for k in "duv":
        dev.set(k + ".src", ref.buf(k + ".src"))
        dev.set(k + ".dst", ref.buf(k + ".dst"))  # non-fluid cells keep stale _dst
    dev.addInflow(*inflow); ref.call("addInflow", *inflow)
    st = dev.update(0.005)
    ref.call("update", 0.005)
    it = re.findall(r"(?:after|of) (\d+) iterations", ref.log())
    return st, int(it[-1])


@pytest.mark.parametrize("version,w,h,steps", [(5, 128, 128, 6), (4, 128, 128, 6), (5, 96, 96, 5)])
def test_update_with_bodies_stepwise(ifl, version, w, h, steps):
    """update() (v5:927-953) one step at a time from the reference's state: bodies at rest as
    in the shipped main() (v5:986).  (With ROTATING bodies the discrete problem becomes
    inconsistent and the reference's own PCG diverges after ~70 iterations -- |r|inf reaches
    4.5e8 inside the 2000-iteration budget at 96^2 -- which leaves nothing to compare; moving
    bodies are covered by the bit-exact per-stage tests above.)"""
    dev, ref, bodies = make_pair(ifl, version, w, h, moving=False)
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    for i in range(steps):
        st, it_ref = one_step_from_reference_state(dev, ref, bodies, inflow)
        assert st[0] == 0, st
        assert abs(st[1] - it_ref) <= max(2, ITER_SLACK * it_ref), "step %d: %d vs %d iterations" % (i, st[1], it_ref)
        assert_bits(dev.get("d.src")[ref.buf("d.cell") != 0], ref.buf("d.src")[ref.buf("d.cell") != 0], "solid d")
        for k in "duv":
            assert rel_err(dev.get(k + ".src"), ref.buf(k + ".src")) <= SOLID_REL, (i, k)
        if i % 4 == 3:
            for b in bodies:
                b.update(0.005)
            ref.call("bodiesUpdate", 0.005)
    dev.close(); ref.close()


# ---- whole trajectories without re-synchronisation, against a MEASURED envelope ---------------------
# Three solvers advance independently from the same start: the device, the reference, and a twin built from
# the unmodified reference's own kernels with nothing but the summation order of its dot products changed
# (tests/envelope.py: update_reordered) -- the liberty the device takes.  What the twin drifts away from the
# reference is the algorithm's own spread at this size and step; the device has to stay within
# envelope.FACTOR times that, and its iteration counts within the twin's spread.
from envelope import Envelope, update_reordered  # noqa: E402


@pytest.mark.parametrize("version,moving", [(4, True), (5, False), (4, False)])
def test_trajectory_within_measured_envelope(ifl, version, moving):
    """10 update() steps at 128^2 without re-synchronising.  Chapter 4 runs with a TRANSLATING body
    (v = (-1, 0.5), bodies advanced every 4th step as in main, v4:960-961): rigid translation keeps the
    binary-cell system solvable.  Chapter 5 keeps its bodies at rest like the shipped main (v5:986): with
    fractional volumes ANY body velocity -- even 1e-3 -- leaves the unmodified reference's PCG at its
    2000-iteration budget without converging (SURVEY's rotating set diverges outright), which leaves
    nothing meaningful to compare."""
    import re
    w = h = 128
    bodies = make_bodies(ifl, moving=False)
    if moving:
        bodies[2].velX, bodies[2].velY = -1.0, 0.5
    rows = [b.as_row() for b in bodies]
    dev = ifl.FluidSolver(w, h, 0.1, version=version, bodies=bodies)
    ref = refapi.Ref(version, w, h, [0.1], rows)
    twin = refapi.Ref(version, w, h, [0.1], rows, fresh_copy=True)  # own copy of the library: own log
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
    envelope = Envelope(floor=10 * REL)  # ten chained solves
    for step in range(10):
        dev.addInflow(*inflow); ref.call("addInflow", *inflow); twin.call("addInflow", *inflow)
        st = dev.update(0.005)
        ref.call("update", 0.005)
        it_twin = update_reordered(twin, 0.005)[-1]
        it_ref = int(re.findall(r"(?:after|of) (\d+) iterations", ref.log())[-1])
        Envelope.check_iterations(st[1], it_ref, it_twin, step)
        for k in "duv":
            e = rel_err(dev.get(k + ".src"), ref.buf(k + ".src"))
            env = rel_err(twin.buf(k + ".src"), ref.buf(k + ".src"))
            print("chapter %d step %d %s: device vs reference %.2e, reference vs re-ordered reference %.2e" % (version, step, k, e, env))
            envelope.check(e, env, (step, k))
        if step % 4 == 3:
            for b in bodies:
                b.update(0.005)
            ref.call("bodiesUpdate", 0.005); twin.call("bodiesUpdate", 0.005)
    print("chapter %d, moving=%s: worst device deviation %.2e at a reference envelope of %.2e" % ((version, moving) + envelope.worst))
    dev.close(); ref.close(); twin.close()


def test_reference_pcg_sensitivity():
    """Pins the reference's own noise floor (CPU only, but kept next to the tolerance it
    justifies): one ulp on one rhs entry moves the reference's converged pressure by far
    more than 1e-10."""
    import importlib
    ifl = importlib.import_module("incremental-fluids_b200")
    w = h = 128
    rows = [b.as_row() for b in make_bodies(ifl, moving=False)]
    out = []
    for flip in (False, True):
        ref = refapi.Ref(5, w, h, [0.1], rows)
        ref.call("addInflow", 0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        for k in "duv":
            ref.call(k + ".fillSolidFields")
        ref.call("setBoundaryCondition"); ref.call("buildRhs"); ref.call("buildPressureMatrix", 0.005)
        ref.call("buildPreconditioner")
        if flip:
            r = ref.buf("r")
            i = int(np.argmax(np.abs(r)))
            r[i] = np.nextafter(r[i], np.inf)
        ref.call("project", 2000)
        out.append(ref.buf("p").copy())
        ref.close()
    noise = rel_err(out[1], out[0])
    assert noise > 1e-10, noise   # the 1e-10 bar is below what the reference reproduces of itself
    assert noise < SOLID_REL, noise
