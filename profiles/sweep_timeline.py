"""Strip timeline of one backward MIC(0) sweep at SIZE^2 (diagnostics for sweep_kernels.cu):
prints per-strip start lag and duration, from which the step time T (cycles per column
step) and the strip-to-strip lag L follow."""
import importlib
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ifl = importlib.import_module("incremental-fluids_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nh = int(sys.argv[2]) if len(sys.argv) > 2 else n
s = ifl.FluidSolver(n, nh, 0.1, version=3)
s.buildPressureMatrix(0.005)
s.buildPreconditioner()
s.set("r", np.random.default_rng(0).uniform(-1, 1, n * nh))
for _ in range(3):
    s.applyPreconditioner("z", "r")
# third argument "dot": time the backward sweep WITH the fused z.r (the variant the PCG loop runs)
if len(sys.argv) > 3 and sys.argv[3] == "dot":
    t = s.sweep_times(lambda: s.project(2)).astype(np.int64)
else:
    t = s.sweep_times(lambda: s.applyPreconditioner("z", "r")).astype(np.int64)
t0 = t[:, 0].min()
start, end = t[:, 0] - t0, t[:, 1] - t0
dur = end - start
steps = n + 31
print("strips %d, columns %d" % (len(t), n))
if len(t) > 1:
    print("first lags (us):", [round(float(x) / 1e3, 2) for x in np.diff(end)[:6]])
print("strip 0 duration %.1f us -> %.1f ns/step" % (dur[0] / 1e3, dur[0] / steps))
print("total %.1f us" % (end.max() / 1e3))
print("strip 0: %.1f SM cycles/step, SM clock during the sweep %.0f MHz" % (t[0, 15] / steps, t[0, 15] / dur[0] * 1e3))
print("end-to-end lag between consecutive strips (us): mean %.2f  min %.2f  max %.2f" %
      (np.diff(end).mean() / 1e3, np.diff(end).min() / 1e3, np.diff(end).max() / 1e3))
print("start lag (us): mean %.2f" % (np.diff(start).mean() / 1e3))
big = [(i + 1, round(float(x) / 1e3, 1)) for i, x in enumerate(np.diff(end)) if x > 2 * np.median(np.diff(end))]
print("lags above twice the median (strip, us):", big[:24])
print("median lag %.2f us" % (np.median(np.diff(end)) / 1e3))
for i in list(range(0, len(t), max(1, len(t) // 16))):
    print("  strip %3d start %8.1f end %8.1f dur %8.1f us" % (i, start[i] / 1e3, end[i] / 1e3, dur[i] / 1e3))
print("checkpoints (us since kernel start) for the first strips:")
for i in range(min(4, len(t))):
    cp = t[i, 2:11] - t0
    print("  strip %d:" % i, [round(float(x) / 1e3, 1) for x in cp])
print("hand-off of the group ending at column %d (us): producer done -> publisher sent -> [next strip] poller released -> compute passed" % ((n // 32 // 2) * 32))
for i in range(1, min(6, len(t))):
    a, b, c, d = t[i - 1, 11], t[i - 1, 12], t[i, 13], t[i, 14]
    print("  strip %d->%d: publish +%.2f  receive +%.2f  consume +%.2f   (total %.2f)" %
          (i - 1, i, (b - a) / 1e3, (c - b) / 1e3 if c else float("nan"), (d - c) / 1e3 if c else float("nan"), (d - a) / 1e3))
