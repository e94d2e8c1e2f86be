#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block of the metrics DESIGN.md quotes per kernel.
usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_summary.py X.csv [extra-metric-substring ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    extra = sys.argv[2:]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== " + r[kn][:110])
        for i, k in enumerate(hdr):
            if k in KEYS or "issue_stalled" in k and k.endswith("per_issue_active.ratio") or any(e in k for e in extra):
                print("  %-86s %-14s %s" % (k, units[i], r[i]))


if __name__ == "__main__":
    main()
