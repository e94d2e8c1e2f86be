#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the handful of metrics DESIGN.md / profiles/ quote.
usage: tools/ncu_summary.py file.ncu-rep [launch_index]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_requests.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    for h, u, v in zip(hdr, units, vals):
        if h == 'Kernel Name':
            print('kernel', v[:100])
        if h in WANT:
            print('%-70s %-14s %s' % (h, u, v))
        elif 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                if float(v.replace(',', '')) > 0.3:
                    print('%-70s %-14s %s' % (h, u, v))
            except ValueError:
                pass


if __name__ == '__main__':
    main()
