#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 1 --warmup 3
--no-cuda-graph` to the launches of the LAST (timed) step and a per-kernel summary.

    python profiles/condense_launches.py gpurun_out/s2_bench_launches_final.csv profiles/s2_bench_launches_final

writes <out>.csv (id, kernel, block, grid, ns) and <out>_summary.txt (time share per kernel name).
The eager bench runs 4 identical steps (3 warm-up + 1 timed): the last quarter of the launches of the step's
first kernel onward is the timed step."""
import collections
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum":
        rows.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Block Size"]], r[ix["Grid Size"]],
                     float(r[ix["Metric Value"]].replace(",", ""))))
# the optimizer kernel ends every step
ends = [i for i, r in enumerate(rows) if "adam_multi_kernel" in r[1]]
assert len(ends) >= 2, "expected one adam_multi_kernel launch per step"
step = rows[ends[-2] + 1:ends[-1] + 1]
short = lambda k: re.sub(r"\(anonymous namespace\)::|csb200::|void |at::native::|<unnamed>::", "", k)[:110]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "block", "grid", "ns"])
    for r in step:
        w.writerow([r[0], short(r[1]), r[2], r[3], int(r[4])])
tot = sum(r[4] for r in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    k = short(r[1]).split("(")[0][:90]
    agg[k][0] += 1
    agg[k][1] += r[4]
ours_ns = sum(r[4] for r in step if "csb200::" in r[1])  # every kernel of libcsb200.so lives in namespace csb200
with open(out + "_summary.txt", "w") as f:
    f.write(f"one eager train step (512^2, batch 32, bf16): {len(step)} launches, {tot / 1e6:.2f} ms summed under ncu "
            f"(cold-cache, serialised: compare shares)\n")
    f.write(f"csb200 kernels: {100 * ours_ns / tot:.1f} % of the summed time, "
            f"{sum(1 for r in step if 'csb200::' in r[1])} of the launches\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        f.write(f"{t / 1e3:9.1f} us {100 * t / tot:5.1f} %  n={n:4d}  {k}\n")
print(len(step), "launches,", round(tot / 1e6, 2), "ms")
