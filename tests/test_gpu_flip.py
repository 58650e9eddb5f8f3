"""GPU parity tests for chapter 8 (FLIP): P2G, G2P, copy/diff/undiff and particle advection
against the UNMODIFIED reference (oracle/_ref/libref_v8.so).  All bit-exact: the device P2G
gathers contributions in ascending particle index, which is the reference's scatter order."""
import math

import numpy as np
import pytest

from oracle import refapi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refapi.available(8), reason="oracle/_ref not built (needs /root/reference)")]

RHO_AIR, RHO_SOOT, DIFFUSION = 0.1, 0.1, 0.01  # v8:1474-1476


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    same = np.array_equal(bits(a), bits(b)) if a.dtype == np.float64 else np.array_equal(a, b)
    if not same:
        bad = np.flatnonzero((bits(a) != bits(b)) if a.dtype == np.float64 else (a != b))
        raise AssertionError("%s: %d of %d differ, first at %d: %r vs %r" %
                             (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]]))


def make_pair(ifl, w, h):
    bodies = [ifl.SolidBox(0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0)]  # v8:1484
    dev = ifl.FluidSolver(w, h, RHO_AIR, version=8, bodies=bodies, rho_soot=RHO_SOOT, diffusion=DIFFUSION)
    ref = refapi.Ref(8, w, h, [RHO_AIR, RHO_SOOT, DIFFUSION], [b.as_row() for b in bodies])
    return dev, ref


def ref_particles(ref):
    n = int(ref.call("qs.particleCount"))
    return n, ref.buf("qs.posX")[:n], ref.buf("qs.posY")[:n], [ref.buf("qs.prop%d" % t)[:n] for t in range(4)]


def scramble(ref, rng, crowd=False):
    """Random properties; optionally pile extra particles into a few cells and shuffle the
    index order (what pruning's swap-with-last does to a real run, v8:776-780)."""
    n, px, py, props = ref_particles(ref)
    for p in props:
        p[:] = rng.uniform(-1.0, 1.0, n)
    if crowd:
        cap = ref.buf("qs.posX").size
        extra = min(2000, cap - n)
        X, Y = ref.buf("qs.posX"), ref.buf("qs.posY")
        X[n:n + extra] = rng.uniform(10.0, 13.0, extra)
        Y[n:n + extra] = rng.uniform(20.0, 22.0, extra)
        for t in range(4):
            ref.buf("qs.prop%d" % t)[n:n + extra] = rng.uniform(-1.0, 1.0, extra)
        n += extra
        perm = rng.permutation(n)
        X[:n] = X[:n][perm]; Y[:n] = Y[:n][perm]
        for t in range(4):
            P = ref.buf("qs.prop%d" % t)
            P[:n] = P[:n][perm]
        ref.call("qs.setParticleCount", n)
    return ref_particles(ref)


@pytest.mark.parametrize("w,h,crowd", [(64, 64, False), (96, 96, True), (100, 100, True)])
def test_from_particles_bit_exact(ifl, w, h, crowd):
    dev, ref = make_pair(ifl, w, h)
    rng = np.random.default_rng(0)
    for k in "dtuv":
        dev.fillSolidFields(k); ref.call(k + ".fillSolidFields")
    n, px, py, props = scramble(ref, rng, crowd)
    dev.setParticles(px, py, props)
    for t, k in enumerate("dtuv"):
        dev.fromParticles(k); ref.call(k + ".fromParticles", t)
        assert_bits(dev.get(k + ".src"), ref.buf(k + ".src"), "fromParticles " + k)
        assert_bits(dev.get_aux(k, "cell"), ref.buf(k + ".cell"), "cell flags " + k)
    dev.close(); ref.close()


def test_grid_to_particles_and_diff_bit_exact(ifl):
    w = h = 80
    dev, ref = make_pair(ifl, w, h)
    rng = np.random.default_rng(1)
    n, px, py, props = scramble(ref, rng)
    dev.setParticles(px, py, props)
    for k in "dtuv":
        a = rng.uniform(-1.0, 1.0, ref.buf(k + ".src").size)
        ref.buf(k + ".src")[:] = a; dev.set(k + ".src", a)
        dev.copy(k); ref.call(k + ".copy")
        b = a + rng.uniform(-0.1, 0.1, a.size)
        ref.buf(k + ".src")[:] = b; dev.set(k + ".src", b)
    alpha = 1e-3  # _flipAlpha v8:1296
    for k in "dtuv":
        dev.diff(k, alpha); ref.call(k + ".diff", alpha)
        assert_bits(dev.get(k + ".src"), ref.buf(k + ".src"), "diff " + k)
    dev.gridToParticles(alpha); ref.call("qs.gridToParticles", alpha)
    _, _, dprops = dev.getParticles()
    for t in range(4):
        assert_bits(dprops[t], ref.buf("qs.prop%d" % t)[:n], "gridToParticles prop %d" % t)
    for k in "dtuv":
        dev.undiff(k, alpha); ref.call(k + ".undiff", alpha)
        assert_bits(dev.get(k + ".src"), ref.buf(k + ".src"), "undiff " + k)
    dev.close(); ref.close()


def test_particles_advect_bit_exact(ifl):
    w = h = 72
    dev, ref = make_pair(ifl, w, h)
    rng = np.random.default_rng(2)
    n, px, py, props = scramble(ref, rng)
    dev.setParticles(px, py, props)
    for k in "uv":  # up to ~4 cells of displacement, some particles leave the domain and get clamped
        a = rng.uniform(-1.0, 1.0, ref.buf(k + ".src").size) * (800.0 / w)
        ref.buf(k + ".src")[:] = a; dev.set(k + ".src", a)
    dev.particlesAdvect(0.005); ref.call("qs.advect", 0.005)
    dx, dy, _ = dev.getParticles()
    assert_bits(dx, ref.buf("qs.posX")[:n], "posX")
    assert_bits(dy, ref.buf("qs.posY")[:n], "posY")
    # and P2G again from the moved particles (re-binning path)
    dev.fromParticles("d"); ref.call("d.fromParticles", 0)
    assert_bits(dev.get("d.src"), ref.buf("d.src"), "fromParticles after advect")
    dev.close(); ref.close()
