#!/usr/bin/env python
"""Tensor-pipe evidence out of an .ncu-rep: the tcgen05 op-path counters (ops executed, % of the pipe's peak), TMEM /
tensor-memory activity and DRAM bytes for the first launches of the report."""
import csv
import subprocess
import sys

path = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, units, data = rows[0], rows[1], rows[2:5]
print("kernel:", [r[h.index("Kernel Name")][:48] for r in data])
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg")
for i, k in enumerate(h):
    vals = [r[i] for r in data]
    hit = k in want or (("utcimma" in k or "utchmma" in k or "utcqmma" in k) and (k.endswith(".sum") or k.endswith("pct_of_peak_sustained_elapsed") or k.endswith("pct_of_peak_sustained_active")) and ".max" not in k and ".min" not in k)
    if hit and any(v not in ("0", "no data", "") for v in vals):
        print(f"{k:95s} {units[i]:10s}", vals)
