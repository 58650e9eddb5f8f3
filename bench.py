#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

A "step" is one FluidSolver::update() (buildRhs, buildPressureMatrix,
buildPreconditioner, MIC(0)-PCG project with the reference's cap of 600 iterations,
applyPressure, 3x advect, flip; v3:433-447) of a double-precision smoke plume with the
shipped chapter-3 constants (v3:470-486) on an S x S grid (default 4096), preceded by
the shipped addInflow call, all through the C ABI of libifl_b200.so.

metric   cell-updates/s = S*S*steps*N / time      (device-resident fields: `value`)
e2e      same metric through the host-buffer path: every step uploads d,u,v from pinned
         host memory, runs the step and downloads d,u,v (timed region includes copies)
roofline per-kernel-class CUDA-event timing inside the timed region (ifl_profile);
         the dominant class' algorithmic bytes / its mean launch time vs the measured
         HBM copy bandwidth in MEASURED_PEAKS.json
cpu_baseline  the reference's own CPU code (oracle/_ref if built, else the C port) on
         one host core, bounded sample, extrapolated per step (stated in `sample`)

--impl reference runs only the CPU reference arm and prints the same JSON line.
N > 1: one process per GPU (torchrun), ONE grid split into row slabs (SURVEY 8e,
csrc/dist.cu): weak scaling, the grid grows with N so that every GPU keeps 4096^2 cells
(side = 4096*sqrt(N) rounded to whole 32-row strips: 4096, 5792, 8192, 11584); the MIC(0)
wavefront and the stencil rows cross slab boundaries over NVLink peer mappings, the
reductions are folded identically on every rank.  Barrier + max-over-ranks timing.
--size S overrides the side at any N (e.g. --size 16384 for the north-star grid).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INFLOW = (0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0)  # v3:485
DT, DENSITY, LIMIT = 0.005, 0.1, 600              # v3:473-474, v3:437

# algorithmic bytes per cell per launch: each array a phase needs is read
# once and each output written once.
ALG_BYTES = {"matvec": 40, "axpy2_norm": 48, "precon_fwd": 40, "precon_bwd": 48, "xpay": 24,
             "advect": 80.0 / 3, "factor": 32, "gs_sweep": 24}


def ncu_traffic(kernel, size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    ncu --set full capture (profiles/ncu_traffic.json), or None if it was taken at another size."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t["grid"] == [size, size]:
            return t["bytes_per_launch"].get(kernel)
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU reference ----
CPU_SAMPLE_SIDE = 4096  # the CPU arm never times a grid larger than this (a 4096^2 sample already takes ~20 s)


def cpu_reference_sample(full_size, cap_iters, full_iters):
    """Times the reference's CPU path on a bounded sample of the workload: assembly +
    project(limit=cap_iters) + applyPressure + 3x advect, one core, on a grid of at most
    CPU_SAMPLE_SIDE^2 cells (every stage is a streaming pass far above the cache sizes, so the
    time per cell does not depend on the grid; larger workloads are scaled by their cell count).
    Returns (seconds per FULL step extrapolated to `full_iters` PCG iterations, detail)."""
    from oracle import refapi
    size = min(full_size, CPU_SAMPLE_SIDE)
    cell_scale = (full_size / float(size)) ** 2
    use_ref = refapi.available(3)
    if use_ref:
        s = refapi.Ref(3, size, size, [DENSITY])
        call = s.call
        ops = {"inflow": lambda: call("addInflow", *INFLOW), "rhs": lambda: call("buildRhs"),
               "matrix": lambda: call("buildPressureMatrix", DT), "precon": lambda: call("buildPreconditioner"),
               "project": lambda n: call("project", n), "pressure": lambda: call("applyPressure", DT),
               "advect": lambda: [call(k + ".advect", DT) for k in "duv"]}
        kind = "reference"
    else:
        from oracle import portapi
        s = portapi.PortSolver(3, size, size, DENSITY)
        ops = {"inflow": lambda: s.addInflow(*INFLOW), "rhs": s.buildRhs, "matrix": lambda: s.buildPressureMatrix(DT),
               "precon": s.buildPreconditioner, "project": lambda n: s.project(n),
               "pressure": lambda: s.applyPressure(DT), "advect": lambda: [s.advect(k, DT) for k in "duv"]}
        kind = "port"
    t = {}

    def timed(name, fn, *a):
        t0 = time.perf_counter()
        fn(*a)
        t[name] = time.perf_counter() - t0

    ops["inflow"]()
    timed("rhs", ops["rhs"])
    timed("matrix", ops["matrix"])
    timed("precon", ops["precon"])
    timed("project0", ops["project"], 0)          # prologue only (one applyPreconditioner, norm, dot)
    ops["rhs"]()
    timed("projectN", ops["project"], cap_iters)  # prologue + cap_iters iterations
    timed("pressure", ops["pressure"])
    timed("advect", ops["advect"])
    per_iter = max(t["projectN"] - t["project0"], 1e-9) / max(cap_iters, 1)
    fixed = t["rhs"] + t["matrix"] + t["precon"] + t["project0"] + t["pressure"] + t["advect"]
    step = (fixed + per_iter * full_iters) * cell_scale
    detail = ("%s CPU code, 1 thread, %dx%d: assembly+prologue+applyPressure+3 advects timed in full (%.2f s), "
              "%d PCG iterations timed (%.3f s/iter), extrapolated to the %d iterations of the device step"
              % ("unmodified reference (oracle/_ref)" if use_ref else "C port of the reference (oracle/ifl_oracle.c)",
                 size, size, fixed, cap_iters, per_iter, full_iters))
    if cell_scale != 1.0:
        detail += "; scaled by %.2f (cells of the %dx%d workload / cells of the sample)" % (cell_scale, full_size, full_size)
    if use_ref:
        s.close()
    return step, kind, detail


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = grid_side(args, args.gpus)
    cap = min(args.cpu_iters, 3)  # ~18 s per sampled step: W + K steps end within a few minutes
    times = []
    kind = detail = None
    for i in range(args.warmup + args.steps):
        step, kind, detail = cpu_reference_sample(size, cap, LIMIT)
        if i >= args.warmup:
            times.append(step)
    mean = float(np.mean(times))
    value = size * size / mean
    line = {"impl": "reference", "metric": "cell-updates/sec (advect+PCG project)", "value": value,
            "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(size, args.gpus),
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": kind, "sample": detail},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def grid_side(args, n):
    """Side of the square grid: 4096 on one GPU; with N GPUs 4096*sqrt(N) rounded to whole
    32-row strips, i.e. a constant 4096^2 cells per GPU (weak scaling)."""
    if args.size:
        return args.size
    return int(round(4096 * n ** 0.5 / 32.0)) * 32


def workload_config(size, n):
    return {"workload": "3-conjugate-gradients smoke plume %dx%d, double, MIC(0)-PCG limit %d, dt %.3f, "
                        "inflow before every update (v3:470-486)" % (size, size, LIMIT, DT),
            "grid": [size, size], "pcg_limit": LIMIT,
            "parallelism": "1 GPU" if n == 1 else
            "%d row slabs of one %dx%d grid, one process per GPU, peer-mapped HBM over NVLink, exact MIC(0) "
            "pipelined across slabs" % (n, size, size),
            "cells_per_gpu": size * size // n,
            "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (17 * size * size * 8 / 1e9 / n)}


# ------------------------------------------------------------------------- ours ----
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ifl = importlib.import_module("incremental-fluids_b200")
    size = grid_side(args, world)
    rdv = "/tmp/ifl_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"))
    s = ifl.FluidSolver(size, size, DENSITY, version=3, device=local, rank=rank, world=world, rendezvous=rdv)
    stream = torch.cuda.ExternalStream(s.stream(), device=local)

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        s.addInflow(*INFLOW)
        return s.update(DT)

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- timed region 1: device-resident steps, per-kernel-class events on
    s.profile(True)
    launches0 = s.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    iters = []
    for _ in range(args.steps):
        iters.append(step()[1])
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = s.launches() - launches0
    prof = s.profile_read()
    s.profile(False)

    # ---- timed region 2: end to end through host buffers (pinned), copies included
    # every rank keeps the rows of its own slab in pinned host memory (one GPU: whole fields)
    host = {k: torch.empty(s.slab_elems(k + ".src"), dtype=torch.float64).pin_memory() for k in "duv"}
    np_host = {k: v.numpy() for k, v in host.items()}
    r0, r1 = s.rows()
    for k in "duv":
        full = s.get(k + ".src")
        w_k = size + 1 if k == "u" else size
        np_host[k][:] = full[r0 * w_k: r0 * w_k + np_host[k].size]
        del full

    def step_e2e():
        for k in "duv":
            s.set_slab(k + ".src", np_host[k])
        s.addInflow(*INFLOW)
        s.update(DT)
        for k in "duv":
            s.get_slab(k + ".src", np_host[k])

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    e1.record(stream)
    barrier()
    ms_e2e_dev = e0.elapsed_time(e1)
    ms_e2e_wall = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e_dev, ms_e2e_wall)  # host-synchronous copies: wall clock is the honest one
    clocks = sampler.summary() if sampler else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        cells = size * size
        value = cells * args.steps / (ms * 1e-3)  # the grid is ONE job over all ranks
        e2e_value = cells * args.steps / (ms_e2e * 1e-3)
        peak, peak_src = measured_peak()
        shares = {k: v[0] for k, v in prof.items() if v[1] > 0}
        total_prof = sum(shares.values())
        alg = dict(ALG_BYTES)
        if "axpy2_norm" not in shares:  # p += alpha s, r -= alpha q, |r|inf ride inside the forward sweep
            alg["precon_fwd"] = ALG_BYTES["precon_fwd"] + ALG_BYTES["axpy2_norm"]
        dom = max((k for k in shares if k in alg), key=lambda k: shares[k])
        dom_ms, dom_n = prof[dom]
        achieved = alg[dom] * cells / world / (dom_ms / dom_n * 1e-3) / 1e9  # per GPU (rank 0's launches)
        # whole-iteration roofline: 200 algorithmic bytes per cell per PCG iteration
        pcg_ms = sum(prof[k][0] for k in ("matvec", "axpy2_norm", "precon_fwd", "precon_bwd", "xpay", "scalar") if k in prof)
        n_iter = prof["matvec"][1]
        iter_gbs = 200.0 * cells * n_iter / (pcg_ms * 1e-3) / 1e9 if n_iter else None
        bytes_io = sum(v.numel() * 8 for v in host.values()) * world
        line = {
            "metric": "cell-updates/sec (advect+PCG project)", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(size, world),
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": bytes_io,
                    "d2h_bytes_per_step": bytes_io, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(dom, size) if world == 1 else None,
                         "algorithmic_bytes_per_launch": alg[dom] * cells / world, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": alg[dom], "mean_launch_ms": dom_ms / dom_n,
                         "share_of_profiled_time": dom_ms / total_prof if total_prof else None},
            "pcg": {"iterations_per_step": iters, "iters_per_s": n_iter / (pcg_ms * 1e-3) if n_iter else None,
                    "algorithmic_gbs_200B_per_cell_iter": iter_gbs,
                    "frac_of_peak": iter_gbs / (peak * world) if iter_gbs else None},
            "kernel_ms": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in prof.items() if v[1] > 0},
        }
        if world == 1 and not args.no_cpu_baseline:
            step_s, kind, detail = cpu_reference_sample(size, args.cpu_iters, int(np.mean(iters)))
            line["cpu_baseline"] = {"value": cells / step_s, "unit": "cell-updates/s", "cores": 1, "kind": kind,
                                    "sample": detail}
        print(json.dumps(line))
    s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="grid side (default: 4096*sqrt(gpus), whole strips)")
    ap.add_argument("--cpu-iters", type=int, default=6, help="PCG iterations timed on the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
