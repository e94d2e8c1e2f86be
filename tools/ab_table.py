#!/usr/bin/env python
"""Tabulate tools/gpu_ab.sh outputs. usage: python tools/ab_table.py gpurun_out/<tag>_<workload>_v*.json"""
import csv, json, sys, os
for f in sys.argv[1:]:
    d = json.load(open(f)); r = d["roofline"]; c = d["config"]; x = r["expansion"]
    print("%-28s step %.3f ms  value %.3e  search %.3f scan %.3f seed %.3f window %.3f  frac %.2f" % (
        os.path.basename(f)[:-5], d["ms_per_step"], d["value"], r["search_kernel"]["launch_ms"], r["scan_ms"],
        x["seed_pass_ms"], x["window_pass_ms"], r["frac"]))
    n = f[:-5] + ".ncu.csv"
    if os.path.exists(n):
        rows = [x for x in csv.reader(open(n)) if len(x) > 10]
        h = rows[0]; ki = h.index("Kernel Name"); mi = h.index("Metric Name"); vi = h.index("Metric Value")
        ker = {}
        for x in rows[1:]:
            ker.setdefault(x[ki][:40], {})[x[mi]] = x[vi]
        for k, m in ker.items():
            g = lambda s: float(m.get(s, "0").replace(",", ""))
            print("    %-40s %7.1f us  dramR %6.1f MB dramW %6.1f MB  L2miss %5.1fM hit %5.1fM  tag %4.1f%% xbar %4.1f%% issue %4.1f%% inst %6.1fM warps %4.1f%%" % (
                k, g("gpu__time_duration.sum") if g("gpu__time_duration.sum") < 1e5 else g("gpu__time_duration.sum") / 1e3, g("dram__bytes_read.sum") if g("dram__bytes_read.sum") < 1e5 else g("dram__bytes_read.sum")/1e6, g("dram__bytes_write.sum") if g("dram__bytes_write.sum") < 1e5 else g("dram__bytes_write.sum")/1e6,
                g("lts__t_sectors_lookup_miss.sum") / 1e6, g("lts__t_sectors_lookup_hit.sum") / 1e6,
                g("lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed"), g("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__inst_executed.sum") / 1e6, g("sm__warps_active.avg.pct_of_peak_sustained_active")))
