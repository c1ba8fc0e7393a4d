#!/usr/bin/env python3
"""Condense an ncu report (read HERE, no GPU needed) into the per-kernel summary CSV kept under profiles/.
usage: ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.csv"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(rep, out):
    # `rep` is an .ncu-rep (read here with ncu) or the raw metric page already exported on the GPU box (ncu -i ... --page raw --csv)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [m for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + cols)
        w.writerow([""] + [units[hdr.index(c)] for c in cols])
        for r in data:
            w.writerow([r[hdr.index("Kernel Name")]] + [r[hdr.index(c)] for c in cols])
    print(f"{out}: {len(data)} kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
