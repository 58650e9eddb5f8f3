"""CPU: the twin the measured-envelope tests rely on (tests/envelope.py).  Sequenced from Python over the
unmodified reference's own methods WITH the reference's own dotProduct, update_reordered must reproduce
FluidSolver::update bit for bit -- so that with numpy's dot product it differs from the reference by the
summation order of the reductions and by nothing else."""
import math
import os
import re

import numpy as np
import pytest

from oracle import refapi
from envelope import update_reordered

HAVE_REF = os.path.exists(os.path.join(refapi.REF_DIR, "libref_v3.so"))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference once: make -C oracle ref)")
@pytest.mark.parametrize("version,params", [(3, [0.1]), (4, [0.1]), (5, [0.1]), (6, [0.1, 0.1, 0.01]), (7, [0.1, 1.0, 0.01])])
def test_resequenced_update_is_the_reference(version, params):
    n = 64
    rows = [[0, 0.5, 0.6, 0.7, 0.1, math.pi * 0.25, 0.0, 0.0, 0.0]] if version >= 4 else []
    ref = refapi.Ref(version, n, n, params, rows)
    twin = refapi.Ref(version, n, n, params, rows, fresh_copy=True)
    inflow = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0) if version < 6 else (0.45, 0.2, 0.1, 0.05, 1.0, ref.call("ambientT") + 300.0, 0.0, 0.0)
    drift = 0.0
    for step in range(3):
        ref.call("addInflow", *inflow)
        twin.call("addInflow", *inflow)
        ref.call("update", 0.005)
        its = update_reordered(twin, 0.005, exact_dot=True)
        printed = [int(x) for x in re.findall(r"(?:after|of) (\d+) iterations", ref.log())]
        assert [i for i in its if i is not None] == printed, (step, its, printed)
        for q in ("dtuv" if version >= 6 else "duv"):
            assert np.array_equal(ref.buf(q + ".src"), twin.buf(q + ".src")), (step, q)
    # and with numpy's summation order it is a different (equally valid) rounding path
    twin2 = refapi.Ref(version, n, n, params, rows, fresh_copy=True)
    twin2.call("addInflow", *inflow)
    ref2 = refapi.Ref(version, n, n, params, rows, fresh_copy=True)
    ref2.call("addInflow", *inflow)
    ref2.call("update", 0.005)
    update_reordered(twin2, 0.005)
    a, b = ref2.buf("u.src"), twin2.buf("u.src")
    drift = float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300))
    assert 0.0 < drift < 1e-6, drift
    for r in (ref, twin, twin2, ref2):
        r.close()
