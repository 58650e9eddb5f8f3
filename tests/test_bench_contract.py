"""bench.py's reference arm (the one leg that runs without a GPU): the JSON line carries the
keys the driver reads, on the workload the device arm is quoted on, and ranks other than 0
stay silent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_line():
    out = run(None, "--size", "96", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "cell-updates/s" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["config"]["grid"] == [96, 96] and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["cores"] == 1 and cb["kind"] in ("reference", "port") and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    out = run({"RANK": "1", "WORLD_SIZE": "2"}, "--size", "64", "--steps", "1", "--warmup", "0", "--gpus", "2")
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_weak_scaling_grid_sides():
    sys.path.insert(0, ROOT)
    import bench

    class A:
        size = 0
    assert [bench.grid_side(A, n) for n in (1, 2, 4, 8)] == [4096, 5792, 8192, 11584]
    A.size = 16384
    assert bench.grid_side(A, 8) == 16384


@pytest.mark.parametrize("config", ["3", "4", "5"])
def test_reference_arm_samples_a_live_solve(config):
    """Chapters 5-8: the sampled PCG iterations must belong to a solve with a non-trivial right-hand side (chapters 6+
    get theirs from the buoyancy of the heat step: without it project() returns at once and the extrapolated CPU time
    collapses -- the sampler therefore replays update() up to the pressure solve, v7:1120-1147, v8:1350-1385)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_v%s.so" % {"3": 5, "4": 7, "5": 8}[config])):
        pytest.skip("oracle/_ref not built")
    out = run(None, "--config", config, "--size", "96", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    # 96^2 cells, >= 10 sampled iterations extrapolated to hundreds: a live solve stays far below 1e6 cell-updates/s
    assert 1e3 < line["value"] < 3e5, line["value"]
    assert "extrapolated to the" in line["cpu_baseline"]["sample"]
