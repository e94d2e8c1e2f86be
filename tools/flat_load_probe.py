#!/usr/bin/env python
"""Load time of an index with and without the flattened-index file (rig_index_save_flat / rig_index_load_flat).

    python tools/flat_load_probe.py --workload c4 [--count-only]

Builds (or takes from .cache/) the workload's logical index, then times: (1) rig_index_create (flatten on the host +
upload), (2) rig_index_save_flat, (3) rig_index_load_flat of that file (read + upload + digest check against the
logical index), and runs a small count batch on both handles to check that they answer identically. One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--count-only", action="store_true", help="smallest locate tables (what ri-count needs)")
    a = ap.parse_args()
    rib = ge.load_package()
    t0 = time.time()
    _, patt, N, m, host = bench.job_inputs(a.workload, 1, 0)
    t_inputs = time.time() - t0
    kw = dict(device=0, phi_jump=1, seed_jump=1) if a.count_only else dict(device=0)
    flat = os.path.join(bench.CACHE, "%s.probe.flat" % a.workload)
    if os.path.exists(flat):
        os.remove(flat)
    t0 = time.time()
    g1 = rib.GpuIndex(host, **kw)
    t_create = time.time() - t0
    t0 = time.time()
    g1.save_flat(flat)
    t_save = time.time() - t0
    size = os.path.getsize(flat)
    t0 = time.time()
    g2 = rib.GpuIndex(host, flat=flat, **kw)
    t_load = time.time() - t0
    assert g2.from_flat
    t_cold = None
    try:   # the same load with the file evicted from the page cache (root only)
        os.sync()
        with open("/proc/sys/vm/drop_caches", "w") as f:
            f.write("3\n")
        t0 = time.time()
        g3 = rib.GpuIndex(host, flat=flat, **kw)
        t_cold = time.time() - t0
        del g3
    except Exception:
        pass
    S = min(N, 20000)
    r1 = g1.count(patt[: S * m], S, m)
    r2 = g2.count(patt[: S * m], S, m)
    same = bool(np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1]))
    os.remove(flat)
    print(json.dumps({"workload": bench.WORKLOADS[a.workload][9], "n": int(host.n), "r": int(host.r), "count_only_tables": a.count_only,
                      "index_device_bytes": int(g1.info.device_bytes), "seed_jump": int(g1.info.seed_jump), "phi_jump": int(g1.info.phi_jump),
                      "inputs_s": round(t_inputs, 1), "create_flatten_upload_s": round(t_create, 2), "save_flat_s": round(t_save, 2),
                      "flat_file_bytes": size, "load_flat_s": round(t_load, 2),
                      "load_flat_cold_page_cache_s": None if t_cold is None else round(t_cold, 2), "answers_identical": same,
                      "note": "load_flat = file read + upload + digest of the logical index; the first figure with the page cache warm from the save"}))


if __name__ == "__main__":
    main()
