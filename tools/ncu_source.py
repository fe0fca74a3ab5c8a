#!/usr/bin/env python
"""Per-instruction stall attribution from an .ncu-rep (source page, SASS): where the warps wait.
    python tools/ncu_source.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio", "stall_lg",
          "stall_not_selected", "stall_selected", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_tex"]


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[start]
    data = []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break            # first kernel only
        if len(r) == len(h):
            data.append(r)
    i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    cols = {k: h.index(k) for k in STALLS if k in h}
    tot = sum(int(r[i_s]) for r in data)
    print("kernel:", rows[start - 1][1][:90] if start else "?")
    print("total samples", tot, "instructions", len(data), "executed warp-inst", sum(int(r[i_ex]) for r in data))
    for k, i in cols.items():
        v = sum(int(r[i]) for r in data)
        if v:
            print(f"  {k:26s} {v:8d} {100.0 * v / max(tot, 1):5.1f}%")
    print("--- top instructions by samples (index, sass, samples, executed, stalls)")
    idx = sorted(range(len(data)), key=lambda j: -int(data[j][i_s]))[:top]
    for j in sorted(idx):
        r = data[j]
        print(j, r[i_src].strip()[:64].ljust(64), r[i_s], r[i_ex], {k.replace("stall_", ""): int(r[i]) for k, i in cols.items() if int(r[i]) > 0})


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
