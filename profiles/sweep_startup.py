import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ifl = importlib.import_module("incremental-fluids_b200")
n = 4096
s = ifl.FluidSolver(n, n, 0.1, version=3)
s.buildPressureMatrix(0.005); s.buildPreconditioner()
s.set("r", np.random.default_rng(0).uniform(-1, 1, n * n))
for _ in range(3): s.applyPreconditioner("z", "r")
t = s.sweep_times(lambda: s.applyPreconditioner("z", "r")).astype(np.int64)
t0 = t[:, 0].min()
for i in (0, 1, 2, 3, 40, 41, 100, 101):
    cp = (t[i, 2:11] - t0) / 1e3
    print("strip %3d start %.2f  m-checkpoints:" % (i, (t[i,0]-t0)/1e3), " ".join("%.2f" % x for x in cp), "| dt:", " ".join("%.2f" % x for x in np.diff(cp)))
print("start-up hand-off (us since kernel start): producer group0 final | publisher sent | consumer poller released | consumer compute saw it | consumer m=0")
for i in (1, 2, 3, 4, 40, 41, 100):
    a, b, c, d, e = t[i-1, 11], t[i-1, 12], t[i, 13], t[i, 14], t[i, 2]
    print("  %3d->%3d: %.2f | +%.2f | +%.2f | +%.2f | +%.2f   (total %.2f)" % (i-1, i, (a-t0)/1e3, (b-a)/1e3, (c-b)/1e3, (d-c)/1e3, (e-d)/1e3, (e-a)/1e3))
