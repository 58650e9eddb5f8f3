"""The C++ drop-in classes (incremental-fluids_b200/host/FluidSolver.hpp) driven by a copy of
the reference's main(): stdout must carry the reference's solver lines with the reference's
iteration counts, and the rendered frames must equal frames rendered from the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "incremental-fluids_b200", "host")


def fnv_bytes(buf):
    h = 0xcbf29ce484222325
    for b in bytes(buf):
        h = ((h ^ b) * 0x100000001b3) & ((1 << 64) - 1)
    return "%016x" % h


def to_image(d):
    shade = ((1.0 - d) * 255.0).astype(np.int64)
    shade = np.clip(shade, 0, 255).astype(np.uint8)
    rgba = np.empty((d.size, 4), dtype=np.uint8)
    rgba[:, :3] = shade[:, None]
    rgba[:, 3] = 0xFF
    return rgba.tobytes()


def test_plume_main_chapter3(port):
    exe = os.path.join(HOST, "plume_v3")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST])
    size, frames = 128, 5
    out = subprocess.run([exe, str(size), str(frames)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    iters = [int(x) for x in re.findall(r"Exiting solver after (\d+) iterations, maximum error is", out.stdout)]
    # SURVEY.md section 4 anchor of the shipped chapter-3 run
    assert iters == [73, 73, 72, 71, 71] + [70] * 15
    hashes = re.findall(r"fnv64\(rgba\)=([0-9a-f]{16})", out.stdout)
    ora = port.PortSolver(3, size, size, 0.1)
    want = []
    for f in range(frames):
        for _ in range(4):
            ora.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)
            ora.update(0.005)
        want.append(fnv_bytes(to_image(ora.src["d"])))
    assert hashes == want
