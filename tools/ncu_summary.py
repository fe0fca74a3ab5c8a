#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers that decide the next optimisation."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path, first=0, count=3):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    data = rows[2 + first:2 + first + count]
    print("kernel:", [r[h.index("Kernel Name")][:50] for r in data], [r[h.index("Grid Size")] for r in data])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:75s} {units[i]:12s}", [r[i] for r in data])
    stall = [x for x in h if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and "average_warps_" in x and "not_issued" not in x]
    print("--- stall reasons (warps per issue-active cycle) ---")
    vals = []
    for k in stall:
        i = h.index(k)
        try:
            vals.append((float(data[0][i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    for v, k in sorted(vals, reverse=True)[:8]:
        print(f"   {k:30s} {v:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 2)
