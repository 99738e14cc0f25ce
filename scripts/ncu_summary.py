#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of per-kernel metrics DESIGN.md / bench.py quote.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt   (runs on the CPU box)"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none summary of {rep}")
    for r in rows[2:]:
        print("----")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k} = {r[i]} {units[i]}")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.3:
                    print(f"{h} = {r[i]}")


if __name__ == "__main__":
    main()
