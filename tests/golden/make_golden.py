"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_v*.so,
i.e. /root/reference/<N>/Fluid.cpp compiled by oracle/Makefile).  Run in the build
container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each file holds the state after `steps` reference updates of a smoke plume with the
shipped constants of that chapter's main() (v3:470-486), at a small grid, plus the
solver status lines the reference printed.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refapi import Ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# main() constants: v1:346-362 (sharper inflow), v2:355-371, v3:470-486
INFLOW = {1: (0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 3.0), 2: (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0),
          3: (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)}
CASES = [(1, 48, 48, 4), (2, 64, 48, 4), (3, 64, 64, 6), (3, 96, 40, 5)]


def main():
    for ver, w, h, steps in CASES:
        r = Ref(ver, w, h, [0.1])
        for _ in range(steps):
            r.call("addInflow", *INFLOW[ver])
            r.call("update", 0.005)
        log = r.log()
        iters = [int(x) for x in re.findall(r"(?:after|of) (\d+) iterations", log)]
        out = os.path.join(HERE, "v%d_%dx%d.npz" % (ver, w, h))
        np.savez_compressed(out, version=ver, w=w, h=h, steps=steps, density=0.1, timestep=0.005,
                            inflow=np.array(INFLOW[ver]), d=r.buf("d.src").copy(), u=r.buf("u.src").copy(),
                            v=r.buf("v.src").copy(), p=r.buf("p").copy(), iters=np.array(iters), log=log)
        print(out, iters)
        r.close()


if __name__ == "__main__":
    main()
