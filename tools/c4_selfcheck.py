#!/usr/bin/env python
"""One-off at-scale check of the 64-bit paths (n >= 2^32): config 4's 10 GB pan-genome, locate + the device -c
self-check (every located position compared with the text, counts against the hash-join brute force), plus the
timing of a locate step on 64-bit words. Builds the index with the prefix-free-parsing builder (about a minute on
the GPU box). usage: python tools/c4_selfcheck.py [N_reads] > gpurun_out/c4_selfcheck.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rib = ge.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
kind, n, p0, p1, tseed, _, m, pseed, limit, desc = bench.WORKLOADS["c4"]
t0 = time.time()
text = rib.gen_text(kind, n, p0, p1, tseed)
patt = rib.gen_patterns(text, N, m, pseed, limit)
path = os.path.join(ROOT, ".cache", "c4.rib")
host = rib.HostIndex.load(path) if os.path.exists(path) else rib.HostIndex.from_text_auto(text)
t1 = time.time()
gpu = rib.GpuIndex(host)
t2 = time.time()
info = gpu.info
gpu.text_attach(text)
del text
lo, hi, off, occ, rep = gpu.locate_ex(patt, N, m, rib.LOCATE_SORT | rib.LOCATE_CHECK)
t3 = time.time()
lo2, hi2, off2, occ2 = gpu.locate(patt, N, m)
tm = gpu.timing()
out = {"workload": desc, "n": int(info.n), "r": int(info.r), "words32": int(info.words32), "phi_jump": int(info.phi_jump),
       "seed_jump": int(info.seed_jump), "index_device_bytes": int(info.device_bytes), "reads": N, "m": m,
       "occurrences": int(occ.size), "check": rep.as_dict(), "clean": bool(rep.clean),
       "ranges_equal": bool(np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(off, off2)),
       "same_multiset": bool(np.array_equal(np.sort(occ2[: int(off[200])]), np.sort(occ[: int(off[200])]))),
       "timing_ms": {k: tm[k] for k in ("search_ms", "scan_ms", "seed_ms", "window_ms", "expand_ms")},
       "prepare_s": round(t1 - t0, 1), "flatten_upload_s": round(t2 - t1, 1), "locate_sort_check_s": round(t3 - t2, 1)}
print(json.dumps(out))
assert out["clean"] and out["ranges_equal"] and out["same_multiset"] and out["words32"] == 0
