"""Turns the ncu --set full reports of profiles/ncu_capture.sh (gpurun_out/ncu_r02/*.ncu-rep) into the committed
evidence: profiles/r02_ncu_kernel_table.md (one row per captured launch) and profiles/ncu_traffic.json
(dram bytes per launch of the kernels bench.py reports a roofline for).  Run in the CPU container:
    python profiles/ncu_summarise.py [gpurun_out/ncu_r02]
"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ncu_r02")

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
# algorithmic bytes per cell of the kernels that have one; the grid is given per report
ALG = {"k_tri<0": 40, "k_tri<1": 48, "k_matvec": 40, "k_xpay_matvec": 64, "k_axpy2_norm": 48, "k_scaled_add": 24, "k_advect": 80.0 / 3,
       "k_sweep<2": 32, "k_sweep<4": 32, "k_sweep<3": 24, "k_build_rhs": 24, "k_build_matrix": 24, "k_apply_pressure": 24}
CELLS = {"gs_sweep": 2048 * 2048, "p2g": 1024 * 1024, "g2p": 1024 * 1024, "padvect": 1024 * 1024}


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def unit_scale(unit, kind):
    u = (unit or "").lower()
    if kind == "time":
        return {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}.get(u, 1.0)
    if kind == "bytes":
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "b": 1.0}.get(u, 1.0)
    return 1.0


rows = []
for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        print("cannot read", rep, out.stderr[:200])
        continue
    rd = list(csv.reader(io.StringIO(out.stdout)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        row = {"report": os.path.basename(rep)[:-8], "kernel": r[col["Kernel Name"]]}
        for m, key in WANT.items():
            if m in col:
                v = to_float(r[col[m]])
                if v is not None:
                    kind = "time" if key == "time_us" else "bytes" if key.startswith("dram_r") or key.startswith("dram_w") else ""
                    v *= unit_scale(units[col[m]], kind)
                row[key] = v
        rows.append(row)

if not rows:
    raise SystemExit("no reports under " + src)

traffic = {"grid": [4096, 4096], "bytes_per_launch": {}, "source": "ncu --set full --clock-control none, profiles/ncu_capture.sh (round 2)"}
lines = ["# ncu --set full, one launch per kernel class (round 2)", "",
         "Captured on one B200 with `profiles/ncu_capture.sh` (`--clock-control none`; replays are cold-cache and serialised, so the",
         "durations are NOT bench values -- bench.py times the same kernels with CUDA events).  `alg` = SURVEY 8d's algorithmic bytes",
         "of the launch; `dram` = dram__bytes_read.sum + dram__bytes_write.sum; `GB/s` = alg / duration.", "",
         "| report | kernel | grid x block | regs | time us | alg MB | dram MB | dram/alg | alg GB/s | dram % peak | warps active % | FP64 pipe % | issue active % | smem bank conflicts |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    k = r["kernel"]
    short = re.sub(r"\(.*", "", k)
    short = re.sub(r"^void ", "", short)
    short = re.sub(r"ifl::((tri|stair)::)?", "", short)
    cells = CELLS.get(r["report"], 4096 * 4096)
    alg = None
    for pat, b in ALG.items():
        if short.startswith(pat.split("<")[0]) and (("<" not in pat) or re.search(re.escape(pat.split("<")[0]) + r"<\(?(bool\))?\(?(int\))?" + pat.split("<")[1], short) or pat in short.replace("(bool)", "").replace("(int)", "")):
            alg = b * cells
            break
    dram = (r.get("dram_read") or 0) + (r.get("dram_write") or 0)
    t = r.get("time_us")
    fmt = lambda v, f="%.1f": (f % v) if v is not None else "-"
    lines.append("| %s | `%s` | %s x %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        r["report"], short[:60], fmt(r.get("grid"), "%d"), fmt(r.get("block"), "%d"), fmt(r.get("regs"), "%d"), fmt(t),
        fmt(alg / 1e6 if alg else None), fmt(dram / 1e6), fmt(dram / alg if alg else None, "%.2f"),
        fmt(alg / (t * 1e-6) / 1e9 if alg and t else None, "%.0f"), fmt(r.get("dram_pct")), fmt(r.get("warps_active_pct")),
        fmt(r.get("fp64_pipe_pct") if r.get("fp64_pipe_pct") is not None else r.get("fp64_inst_pct")), fmt(r.get("issue_active_pct")),
        fmt(r.get("smem_bank_conflicts"), "%d")))
    name = short.replace("(bool)", "").replace("(int)", "").replace(" ", "")
    if r["report"] in ("tri", "matvec", "xpay_matvec", "axpy2_norm", "scaled_add"):
        traffic["bytes_per_launch"][name] = dram
with open(os.path.join(ROOT, "profiles", "r02_ncu_kernel_table.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print("\n".join(lines))
print(json.dumps(traffic, indent=1))
