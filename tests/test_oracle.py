"""CPU tests of the oracle (the C restatement, oracle/ifl_oracle.c):
  * against the golden vectors generated from the unmodified reference (always);
  * live against the unmodified reference, oracle/_ref/libref_v*.so (where it was built);
  * the survey's reproducible anchors for the shipped 128^2 run.
"""
import glob
import os
import re

import numpy as np
import pytest

from oracle import refapi

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_reproduces_golden(port, path):
    g = np.load(path)
    s = port.PortSolver(int(g["version"]), int(g["w"]), int(g["h"]), float(g["density"]))
    iters = []
    for _ in range(int(g["steps"])):
        s.addInflow(*g["inflow"])
        st = s.update(float(g["timestep"]))
        iters.append(st[1])
    assert iters == list(g["iters"])
    for k in "duv":
        assert np.array_equal(bits(s.src[k]), bits(g[k])), k
    assert np.array_equal(bits(s.p), bits(g["p"]))


needs_ref = pytest.mark.skipif(not refapi.available(3), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("ver,w,h", [(1, 40, 40), (2, 56, 40), (3, 72, 56), (3, 37, 64)])
def test_port_matches_reference_trajectory(port, ver, w, h):
    r = refapi.Ref(ver, w, h, [0.1])
    p = port.PortSolver(ver, w, h, 0.1)
    inflow = (0.3, 0.2, 0.3, 0.1, 1.0, 0.5, 3.0)
    for _ in range(4):
        r.call("addInflow", *inflow)
        p.addInflow(*inflow)
        r.call("update", 0.005)
        p.update(0.005)
    for k in "duv":
        assert np.array_equal(bits(r.buf(k + ".src")), bits(p.src[k])), k
    assert np.array_equal(bits(r.buf("p")), bits(p.p))
    r.close()


@needs_ref
def test_port_matches_reference_granular(port):
    """Each PCG helper on random data, bit for bit (v3:208-398)."""
    w, h = 45, 38
    rng = np.random.default_rng(0)
    r = refapi.Ref(3, w, h, [0.1])
    p = port.PortSolver(3, w, h, 0.1)
    for k in "uv":
        a = rng.uniform(-1, 1, p.src[k].size)
        r.buf(k + ".src")[:] = a
        p.src[k][:] = a
    r.call("buildRhs"); p.buildRhs()
    r.call("buildPressureMatrix", 0.005); p.buildPressureMatrix(0.005)
    r.call("buildPreconditioner"); p.buildPreconditioner()
    for name, arr in (("r", p.r), ("aDiag", p.aDiag), ("aPlusX", p.aPlusX), ("aPlusY", p.aPlusY), ("precon", p.precon)):
        assert np.array_equal(bits(r.buf(name)), bits(arr)), name
    s = rng.uniform(-1, 1, w * h)
    r.buf("s")[:] = s; p.s[:] = s
    r.call("matrixVectorProduct", 2, 3); p.matrixVectorProduct(p.z, p.s)
    assert np.array_equal(bits(r.buf("z")), bits(p.z))
    r.call("applyPreconditioner", 2, 0); p.applyPreconditioner(p.z, p.r)
    assert np.array_equal(bits(r.buf("z")), bits(p.z))
    assert r.call("dotProduct", 2, 0) == p.dotProduct(p.z, p.r)
    assert r.call("infinityNorm", 0) == p.infinityNorm(p.r)
    r.call("project", 50); st = p.project(50)
    assert np.array_equal(bits(r.buf("p")), bits(p.p))
    assert ("after %d iterations" % st[1]) in r.log() or st[0] != 0
    r.close()


@needs_ref
def test_reference_anchor_v3_128():
    """SURVEY.md section 4 anchors for the shipped chapter-3 run: iteration counts of the
    first 20 updates and sum(d) after 100 updates (493.71060526123603)."""
    r = refapi.Ref(3, 128, 128, [0.1])
    for i in range(100):
        r.call("addInflow", 0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
        r.call("update", 0.005)
    iters = [int(x) for x in re.findall(r"after (\d+) iterations", r.log())]
    assert iters[:20] == [73, 73, 72, 71, 71] + [70] * 15
    total = 0.0
    for x in r.buf("d.src").tolist():
        total += x
    assert total == 493.71060526123603
    r.close()
