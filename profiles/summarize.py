#!/usr/bin/env python
"""Extract the judged numbers from ncu reports (gpurun_out/*.ncu-rep) into profiles/*.csv.

    python profiles/summarize.py gpurun_out/r1_simam_nchw_fwd_bf16.ncu-rep [...]

Writes <name>.metrics.csv (selected raw metrics, one row per captured launch) and, when the report
has source info, <name>.hot.csv (top-30 SASS lines by stall samples)."""
import csv
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
OUT = os.path.dirname(os.path.abspath(__file__))


def ncu(rep, page):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                                          text=True).stdout.splitlines()))


for rep in sys.argv[1:]:
    name = os.path.basename(rep).replace(".ncu-rep", "")
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in WANT or h == "Kernel Name"]
    with open(os.path.join(OUT, name + ".metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i][:90] for i in cols])
    src = ncu(rep, "source")
    if len(src) > 3:
        h = src[1]
        ix = {k: i for i, k in enumerate(h)}
        if "# Samples" in ix:
            n = len(h)
            body = sorted([r for r in src[2:] if len(r) == n and (r[ix["# Samples"]] or "0").isdigit()], key=lambda r: -int(r[ix["# Samples"]] or 0))[:30]
            with open(os.path.join(OUT, name + ".hot.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["samples", "instructions_executed", "sass"])
                for r in body:
                    w.writerow([r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:120]])
    print("wrote", name)
