#!/usr/bin/env python
"""Top sampled SASS instructions of an `ncu --page source --csv` export.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > K.csv ; python tools/ncu_hot.py K.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ci = {k: i for i, k in enumerate(hdr)}
S = ci["# Samples"] if "# Samples" in ci else ci["Warp Stall Sampling (All Samples)"]
tot = sum(int(r[S] or 0) for r in body)
print("instructions", len(body), "samples", tot, "warp-instr executed", sum(int(r[ci["Instructions Executed"]] or 0) for r in body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
for i in sorted(order):
    r = body[i]
    print("%4d %6.2f%% exec=%-9s thr=%-5s %s" % (i, 100.0 * int(r[S] or 0) / max(tot, 1), r[ci["Instructions Executed"]],
                                               r[ci["Avg. Threads Executed"]][:5], r[ci["Source"]][:110]))
